"""Runs the s/t nets of one config-2 28x28x64 channel coupling layer (11 launches) a few times.
Used under ncu:  python tools/profile_layer.py [B] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mask = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
preset = sys.argv[4] if len(sys.argv) > 4 else "big"
PRESETS = {
    "big": ([28, 28, 2], 8, 64, [1, 2, 4]),            # config 2, level 0: 28x28x64 channel layer
    "small": ([14, 14, 4], 4, 32, [1, 2]),             # config 2, level 1: io (14,14,4), nk 32, cardinality 4
    "cfg4": ([64, 64, 6], 4, 64, [1, 2, 4, 8]),        # config 4 (light), level 0: groups of 16 / 8 / 4 / 2
    "cfg5": ([128, 128, 4], 2, 64, [1, 2, 4, 8, 16]),  # config 5 (light), level 0: groups of 32 / 16 / 8 / 4 / 2
}
shape, card, nk, dil = PRESETS[preset]
layer = coupling_layer(shape, mask, 3, card, nk, 3, None, LAYER_NORM=True, which_dilations=dil, device="cuda:0")
u = torch.randn(B, *shape, device="cuda:0")
for _ in range(reps):
    v, s, _ = layer.forward_and_Jacobian(u, 0.0, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    v, s, _ = layer.forward_and_Jacobian(u, 0.0, None)
e1.record()
torch.cuda.synchronize()
print(f"layer fwd: {e0.elapsed_time(e1) / reps:.3f} ms (B={B}, mask={mask})")
