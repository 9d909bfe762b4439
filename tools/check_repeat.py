"""Run-to-run reproducibility of one coupling layer's s/t nets: same input twice -> bitwise equal outputs?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
torch.manual_seed(0)
for (shape, mask, nk, card, dil, B) in [([28, 28, 2], 2, 64, 8, [1, 2, 4], 2), ([28, 28, 2], 2, 64, 8, [1, 2, 4], 37),
                                         ([28, 28, 2], 0, 64, 8, [1, 2, 4], 2), ([14, 14, 4], 2, 32, 4, [1, 2], 5),
                                         ([14, 14, 4], 0, 32, 4, [1, 2], 3)]:
    layer = coupling_layer(shape, mask, 3, card, nk, 3, None, LAYER_NORM=True, which_dilations=dil, device="cuda:0")
    info = layer._info
    u = torch.randn(B, info.h, info.w, info.c1, device="cuda:0")
    outs = []
    for rep in range(4):
        A = layer.A_wrapper(u).clone()
        b = layer.b_wrapper(u).clone()
        outs.append((A, b))
    for rep in range(1, 4):
        dA = (outs[rep][0] - outs[0][0]).abs()
        db = (outs[rep][1] - outs[0][1]).abs()
        print(shape, mask, B, "rep", rep, "A maxdiff %.3e (n=%d)" % (dA.max().item(), (dA > 0).sum().item()),
              "b maxdiff %.3e (n=%d)" % (db.max().item(), (db > 0).sum().item()), "| |b|max %.3e" % outs[0][1].abs().max().item())

# batch independence and agreement with the per-branch kernels (set CNF_GC_OCT=0 in a second process to compare files)
layer = coupling_layer([28, 28, 2], 2, 3, 8, 64, 3, None, LAYER_NORM=True, which_dilations=[1, 2, 4], device="cuda:0")
info = layer._info
u = torch.randn(256, info.h, info.w, info.c1, device="cuda:0")
A256 = layer.A_wrapper(u).clone()
for n in (1, 2, 3, 7, 8):
    An = layer.A_wrapper(u[:n].contiguous()).clone()
    d = (An - A256[:n]).abs()
    print("sub-batch", n, "maxdiff %.3e" % d.max().item(), "per-sample", [float("%.2e" % d[i].max().item()) for i in range(n)])
tag = os.environ.get("CNF_GC_OCT", "1")
torch.save(A256.cpu(), f"gpurun_out/A256_oct{tag}.pt")
if os.path.exists("gpurun_out/A256_oct0.pt") and os.path.exists("gpurun_out/A256_oct1.pt"):
    a0, a1 = torch.load("gpurun_out/A256_oct0.pt"), torch.load("gpurun_out/A256_oct1.pt")
    d = (a0 - a1).abs()
    print("oct vs per-branch: maxdiff %.3e, |A|max %.3e; per-sample max (first 8):" % (d.max().item(), a0.abs().max().item()),
          [float("%.2e" % d[i].max().item()) for i in range(8)])
