"""clock64() stamps of the fused grouped-conv kernel (CNF_OCT_DBG=128): thread 0 of the first CTA of every octet (net 0),
per item: top, copies landed, after barrier A, after transform, after barrier B, after the branches, after barrier C."""
import os, sys, ctypes
os.environ['CNF_OCT_DBG'] = os.environ.get('CNF_OCT_DBG_EXTRA', '128')   # add 8: no epilogue, 32: no weight loads
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arl_conditional_normalizing_flows_b200 import _lib
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
B = 256
layer = coupling_layer([28, 28, 2], 2, 3, 8, 64, 3, None, LAYER_NORM=True, which_dilations=[1, 2, 4], device="cuda:0")
layer.A_wrapper(torch.randn(B, 28, 28, 1, device="cuda:0"))
ws = layer._workspace(B)
br = _lib.Borrowed()
for _ in range(2):
    _lib.check(_lib.lib.cnf_debug_pw_conv(layer._h, br(layer.params), br(ws), B, 2, _lib.stream_ptr()))
torch.cuda.synchronize()
buf = (ctypes.c_int64 * 8192)()
_lib.check(_lib.lib.cnf_debug_read_clocks(buf, 8192))
c = np.array(buf[:2048], dtype=np.int64).reshape(8, 32, 8)
names = ["wait copies", "barrier A", "transform", "barrier B", "branches", "barrier C", "flush+loop"]
for o in (0, 2, 4, 7):
    t = c[o]
    n = int((t[:, 0] > 0).sum())
    print(f"octet {o}: {n} items; cycles per phase (items 1..):")
    for it in range(1, min(n, 6)):
        d = np.diff(t[it, :7]).tolist() + [int(t[it + 1, 0] - t[it, 6]) if it + 1 < n else 0]
        print("   item", it, dict(zip(names, d)), "total", int(t[it + 1, 0] - t[it, 0]) if it + 1 < n else None)
    print("   kernel span for this CTA:", int(t[n - 1, 6] - t[0, 0]), "cycles; first stamp offset vs octet 0:", int(t[0, 0] - c[0, 0, 0]))
