"""Times the dominant 1x1-conv kernel (pw_tc3_kernel<64>) alone on the workspace of a config-2 28x28x64 channel layer:
which=0: 64->64 (X -> Y1), which=1: 112->64 + residual.  Same measurement as bench.py's `roofline`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arl_conditional_normalizing_flows_b200 import _lib
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = coupling_layer([28, 28, 2], 2, 3, 8, 64, 3, None, LAYER_NORM=True, which_dilations=[1, 2, 4], device=dev)
info = layer._info
hw, nk, cat = info.h * info.w, info.nk, info.cat
layer.A_wrapper(torch.randn(B, info.h, info.w, info.c1, device=dev))
ws = layer._workspace(B)
tot_b = tot_t = 0.0
for which, nbytes in ((0, 2 * B * hw * (nk + nk) * 4), (1, 2 * B * hw * (cat + nk + nk) * 4)):
    br = _lib.Borrowed()
    pp, pw_ = br(layer.params), br(ws)
    for _ in range(3):
        _lib.check(_lib.lib.cnf_measure_stage(layer._h, pp, pw_, B, which, _lib.stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(_lib.lib.cnf_measure_stage(layer._h, pp, pw_, B, which, _lib.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot_b += nbytes; tot_t += ms
    print(f"which={which}: {ms * 1e3:.1f} us  {nbytes / ms / 1e6:.0f} GB/s algorithmic")
print(f"both: {tot_b / tot_t / 1e6:.0f} GB/s = {tot_b / tot_t / 1e6 / 6552:.3f} of 6552")
