"""Times the stand-alone fused coupling-law kernel (12 B per element of u) on a [512,128,128,4] tensor, masks 0-3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arl_conditional_normalizing_flows_b200 import _lib
dev = torch.device('cuda:0')
B, H, W, D = 512, 128, 128, 4
u = torch.randn(B, H, W, D, device=dev)
v = torch.empty_like(u)
ld = torch.empty(B, device=dev)
for m in (2, 3, 0, 1):
    shp = (B, H, W, D // 2) if m >= 2 else (B, H // 2, W // 2, 2 * D)
    s = 0.1 * torch.randn(*shp, device=dev)
    t = torch.randn(*shp, device=dev)
    def law():
        br = _lib.Borrowed()
        _lib.check(_lib.lib.cnf_coupling_law(br(u), br(s), br(t), m, 0, br(v), br(ld), _lib.stream_ptr()))
    for _ in range(3):
        law()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        law()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"mask {m}: {ms * 1e3:.1f} us  {12 * u.numel() / ms / 1e6:.0f} GB/s")
