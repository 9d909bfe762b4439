"""Per-tensor gradient errors of the CUDA backward vs the fp64 autograd oracle (and torch fp32 autograd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle.flow_torch import FlowOracle
from oracle.grad_torch import loss_and_grads
from oracle.weights import init_weights, synth_inputs
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow

CFG2_R1 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[1] * 4,
               num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
ONE = dict(io_shape=[14, 14, 4], x_d=2, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
           num_kernels_list=[32], cardinality_list=[4])
kind = sys.argv[1] if len(sys.argv) > 1 else 'init'
layers = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 and sys.argv[2] != '-' else None
cfgname = sys.argv[3] if len(sys.argv) > 3 else 'cfg2'
B = int(sys.argv[4]) if len(sys.argv) > 4 else 3
CFG2_FULL = dict(CFG2_R1, ResNeXt_block_list=[3] * 4)
ONE3 = dict(ONE, ResNeXt_block_list=[3])
ONE2 = dict(ONE, ResNeXt_block_list=[2])
cfg = {'cfg2': CFG2_R1, 'cfg2full': CFG2_FULL, 'one': ONE, 'one3': ONE3, 'one2': ONE2}[cfgname]
lam = float(os.environ.get('LAMBDA_Y', 100))
m = cFlow(**cfg, lambda_y=lam, device='cuda:0')
o = FlowOracle(**cfg, lambda_y=lam, dtype=torch.float64)
W = init_weights(o.plan, kind, seed=1)
o.set_weights(W); m.set_weights(W)
xy = synth_inputs('cfg2', B, seed=0) if cfgname.startswith('cfg2') else synth_inputs('noise:14x14x4', B, seed=0)
if os.environ.get('PERTURB'):
    xy = (xy + float(os.environ['PERTURB']) * np.random.default_rng(9).standard_normal(xy.shape)).astype(np.float32)
f64, g64 = loss_and_grads(o, xy.astype(np.float64))
f32, g32 = loss_and_grads(o, xy, dtype=torch.float32)
four, _ = m.loss_and_grad(torch.from_numpy(xy).cuda())
print('loss', [float(t) for t in four], f64)
got = m.grad_views()
rows = []
for li in range(len(g64)):
    for net in 'Ab':
        for k, ref in g64[li][net].items():
            a = got[li][net][k].cpu().numpy().astype(np.float64).reshape(np.shape(ref))
            b = np.asarray(g32[li][net][k], np.float64).reshape(np.shape(ref))
            s = max(np.abs(ref).max(), 1e-30)
            rows.append((np.abs(a - ref).max() / s, np.abs(b - ref).max() / s, li, net, k, s))
import collections
per = collections.defaultdict(float)
for r in rows:
    per[(r[2], r[3])] = max(per[(r[2], r[3])], r[0])
print('max gpu_err per (layer, net):', {k: float('%.1e' % v) for k, v in sorted(per.items())})
rows.sort(reverse=True)
print('tensors with gpu_err > 2e-3:', sum(r[0] > 2e-3 for r in rows), ' with fp32 autograd err > 2e-3:', sum(r[1] > 2e-3 for r in rows), 'of', len(rows))
for r in rows[:25]:
    print("gpu_err %.2e  fp32_err %.2e  L%d %s %-22s scale %.3e" % r)
if layers:
    for r in sorted(rows, key=lambda r: (r[2], r[3], r[4])):
        if r[2] in layers:
            print("gpu_err %.2e  fp32_err %.2e  L%d %s %-22s scale %.3e" % r)
if os.environ.get('CNF_DUMP'):
    torch.cuda.synchronize()
    np.save(os.environ['CNF_DUMP'] + '_ws.npy', m._train_ws.cpu().numpy().view(np.float32))
    np.save(os.environ['CNF_DUMP'] + '_grads.npy', m._grads.cpu().numpy())
