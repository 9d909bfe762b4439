"""clock64 stamps of CTA 0 of pw_tc3_kernel (debug library: make -C .../csrc debug; CNF_PW_DBG must include 128).
usage: CNF_PW_DBG=128 python tools/tc3_clocks.py [which=0|1] [B]   -> per-chunk cycle counts of the transform / MMA / epilogue roles"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from arl_conditional_normalizing_flows_b200 import _lib
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = coupling_layer([28, 28, 2], 2, 3, 8, 64, 3, None, LAYER_NORM=True, which_dilations=[1, 2, 4], device=dev)
info = layer._info
layer.A_wrapper(torch.randn(B, info.h, info.w, info.c1, device=dev))
ws = layer._workspace(B)
br = _lib.Borrowed()
pp, pw_ = br(layer.params), br(ws)
for _ in range(3):
    _lib.check(_lib.lib.cnf_measure_stage(layer._h, pp, pw_, B, which, _lib.stream_ptr()))
torch.cuda.synchronize()
raw = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * 8192)()
raw.cnf_debug_read_clocks.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
assert raw.cnf_debug_read_clocks(buf, 8192) == 0
c = np.array(buf[:], dtype=np.int64).reshape(-1, 64, 8)      # [role][chunk or tile][slot]
t0 = c[0, 0, 0]
np.set_printoptions(linewidth=200)
print("transform role: chunk, loop top (abs), then deltas: raw_full wait, bar_free wait, transform, fence+syncwarp, arrives; chunk period")
for gi in range(4, 28):
    r = c[0, gi]
    print(gi, r[0] - t0, r[3] - r[0], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[7] - r[6], "| period", c[0, gi + 1, 0] - r[0])
print("MMA role: chunk, loop top (abs), bar_full wait, MMA issue, commits; period")
for gi in range(4, 28):
    r = c[1, gi]
    print(gi, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], "| period", c[1, gi + 1, 0] - r[0])
print("epilogue role (warp TW+1): tile, before tfull wait (abs), tfull wait, ld+release, bias/stats/stores; period")
for tl in range(2, 12):
    r = c[2, tl]
    print(tl, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], "| period", c[2, tl + 1, 0] - r[0])
print("producer role (thread 0): chunk, loop top (abs), raw_free wait, copies issued; period")
for gi in range(4, 28):
    r = c[3, gi]
    print(gi, r[0] - t0, r[1] - r[0], r[2] - r[1], "| period", c[3, gi + 1, 0] - r[0])
