"""Prints GPU-vs-fp64-oracle relative errors on config 2 ('rand' and 'init' weights)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
from oracle.flow_torch import FlowOracle
from oracle.weights import init_weights, synth_inputs
CFG2 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())
for kind in ('rand', 'init'):
    for seed in (2, 7):
        o = FlowOracle(**CFG2, dtype=torch.float64)
        W = init_weights(o.plan, kind, seed=seed)
        o.set_weights(W)
        m = cFlow(**CFG2, device="cuda:0"); m.set_weights(W)
        x = synth_inputs('cfg2', 2, seed=3)
        four, ps = o.log_loss(x.astype(np.float64))
        got = [float(t) for t in m.log_loss(torch.from_numpy(x).cuda())]
        z = synth_inputs('noise:28x28x2', 2, seed=9); z[..., 1:] = x[..., 1:]
        xs = o.call(z.astype(np.float64), -1)
        gs = m(torch.from_numpy(z).cuda(), -1).cpu().numpy()
        ld = m.last_per_sample['logdet'].cpu().numpy()
        print(f"{kind} seed{seed}: zy {rel(m.last_per_sample['zy'].cpu().numpy(), ps['zy']):.2e}  logdet_ps {np.abs(ld-ps['logdet']).max()/np.abs(ps['logdet']).max():.2e}  "
              f"ll_z {np.abs(m.last_per_sample['ll_z'].cpu().numpy()-ps['ll_z']).max()/np.abs(ps['ll_z']).max():.2e}  loss {abs(got[0]-four[0])/abs(four[0]):.2e}  sample {rel(gs, xs):.2e}")
