"""Times log_loss + sampling of one BASELINE config (default cfg3: 32x32x4) at batch B on cuda:0 (CUDA events).
usage: python tools/time_config.py [cfg3|cfg4|cfg5] [B] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
CFGS = {
    'cfg2': dict(io_shape=[28, 28, 2], x_d=1, num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4]),
    'cfg3': dict(io_shape=[32, 32, 4], x_d=3, num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4]),
    'cfg4': dict(io_shape=[64, 64, 6], x_d=3, num_kernels_list=[64, 64, 32, 32], cardinality_list=[4, 4, 2, 2]),
    'cfg5': dict(io_shape=[128, 128, 4], x_d=3, num_kernels_list=[64, 64, 32, 32], cardinality_list=[2, 2, 2, 2]),
}
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = dict(CFGS[name], squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4)
m = cFlow(**cfg, device="cuda:0")
H, W, D = cfg['io_shape']
x = torch.rand(B, H, W, D, device="cuda:0")
z = torch.randn(B, H, W, D, device="cuda:0")
for _ in range(2):
    m.log_loss(x); m(z, -1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    four = m.log_loss(x)
    xs = m(z, -1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{name} B={B}: {ms:.2f} ms per (log_loss + sampling) = {2 * B / ms * 1e3:.0f} images/s; loss {float(four[0]):.4g}, "
      f"bits/dim {m.bits_per_dim(four[1], four[3]):.4f}, samples finite: {bool(torch.isfinite(xs).all())}")
