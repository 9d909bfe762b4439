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
small = len(sys.argv) > 4 and sys.argv[4] == "small"      # a level-1 layer of config 2: io (14,14,4), nk 32, cardinality 4
shape, card, nk, dil = ([14, 14, 4], 4, 32, [1, 2]) if small else ([28, 28, 2], 8, 64, [1, 2, 4])
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
