"""Dumps the clock64() stamps of pw_tc3_kernel's CTA 0 (CNF_PW_DBG must include 128): per chunk, cycles between the
hand-off points of a transform warp, the MMA issuer and an epilogue warp."""
import os, sys, ctypes
os.environ.setdefault('CNF_PW_DBG', '128')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arl_conditional_normalizing_flows_b200 import _lib
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
B = 256
dev = torch.device('cuda:0')
torch.manual_seed(0)
layer = coupling_layer([28, 28, 2], 2, 3, 8, 64, 3, None, LAYER_NORM=True, which_dilations=[1, 2, 4], device=dev)
layer.A_wrapper(torch.randn(B, 28, 28, 1, device=dev))
ws = layer._workspace(B)
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
br = _lib.Borrowed()
for _ in range(3):
    _lib.check(_lib.lib.cnf_debug_pw_conv(layer._h, br(layer.params), br(ws), B, which, _lib.stream_ptr()))
torch.cuda.synchronize()
buf = (ctypes.c_int64 * 8192)()
_lib.check(_lib.lib.cnf_debug_read_clocks(buf, 8192))
c = np.array(buf[:], dtype=np.int64).reshape(-1, 64, 8)
t0 = c[0, 0, 0]
tr, mm, ep = c[0] - t0, c[1] - t0, c[2] - t0
print("transform warp 0 (cycles since its first stamp): top, after cp.async wait, after bar.sync, after issue, after wait(free), "
      "after body, after fence+syncwarp, after arrive")
for gi in range(4, 24):
    print(gi, tr[gi].tolist(), ' d:', np.diff(tr[gi]).tolist(), ' chunk:', int(tr[gi + 1, 0] - tr[gi, 0]))
print("MMA issuer: before wait(full), after wait, after MMA issue, after commits")
for gi in range(4, 24):
    print(gi, mm[gi, :4].tolist(), ' d:', np.diff(mm[gi, :4]).tolist())
print("epilogue warp: before wait(tfull), after wait, after arrive(tempty)")
for tl in range(2, 12):
    print(tl, ep[tl, :3].tolist(), ' d:', np.diff(ep[tl, :3]).tolist())
print("first/last transform stamps:", int(tr[0, 0]), int(tr[43, 7]) if tr[43, 7] > 0 else None)
for gi in list(range(0, 5)) + list(range(36, 44)):
    if gi % 2 == 0: print("   tile prologue stamps (after cf, after load_stats):", int(tr[gi,1]), int(tr[gi,2]), " prev chunk end", int(tr[gi-1,7]) if gi else 0)
    print(gi, [int(tr[gi, k]) for k in (0, 3, 4, 5, 6, 7)], 'mma', mm[gi, :4].tolist())
print('epi', [ep[tl, :3].tolist() for tl in list(range(0, 3)) + list(range(18, 22))])
