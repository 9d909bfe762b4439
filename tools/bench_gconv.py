"""Times the grouped-conv stage (gconv_oct_kernel, all dilation branches) of a config-2 28x28x64 channel layer alone, the same
measurement as bench.py's `roofline` (cnf_measure_stage which = 2).  usage: python tools/bench_gconv.py [B] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arl_conditional_normalizing_flows_b200 import _lib
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda:0")
torch.manual_seed(0)
layer = coupling_layer([28, 28, 2], 2, 3, 8, 64, 3, None, LAYER_NORM=True, which_dilations=[1, 2, 4], device=dev)
layer.A_wrapper(torch.randn(B, 28, 28, 1, device=dev))
ws = layer._workspace(B)
br = _lib.Borrowed()
pp, pw_ = br(layer.params), br(ws)
for _ in range(3):
    _lib.check(_lib.lib.cnf_measure_stage(layer._h, pp, pw_, B, 2, _lib.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    _lib.check(_lib.lib.cnf_measure_stage(layer._h, pp, pw_, B, 2, _lib.stream_ptr()))
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
flops = 2 * 2 * B * 784 * 9 * (64 * 8 + 32 * 4 + 16 * 2)
print(f"gconv_oct: {us:.1f} us  {flops / us / 1e6:.1f} TFLOP/s fp32 = {flops / us / 1e6 / 74.45:.3f} of the FFMA peak")
