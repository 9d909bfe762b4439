"""Generates tests/golden/*.npz from the oracle (fp64).  Run from the repo root:

    python -m oracle.make_golden

These vectors are the ORACLE'S OWN outputs (regression fixtures); the vectors that pin the oracle to the reference's source
are tests/golden/refsrc_* (oracle/make_ref_golden.py).  The reference itself (TensorFlow) cannot run here, so these are outputs of the
restatement, committed so that (i) the oracle cannot drift silently and (ii) the GPU box, which has no
/root/reference, checks the CUDA path against fixed numbers.  Each file stores the config, the full
weight set, inputs and fp64 outputs of cFlow.call(+1), log_loss and cFlow.call(-1).
"""
import json
import os

import numpy as np
import torch

from .flow_torch import FlowOracle
from .toy import ToyOracle, toy_init_weights
from .weights import init_weights, synth_inputs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CONFIGS = {
    "tiny": (dict(io_shape=[4, 4, 2], x_d=1, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
                  num_kernels_list=[8], cardinality_list=[2]), 3),
    "small_sq": (dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
                      num_kernels_list=[16, 8], cardinality_list=[2, 2]), 2),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (cfg, B) in CONFIGS.items():
        o = FlowOracle(**cfg, dtype=torch.float64)
        W = init_weights(o.plan, 'rand', seed=42)
        o.set_weights(W)
        H, Wd, D = cfg['io_shape']
        x = synth_inputs(f'noise:{H}x{Wd}x{D}', B, seed=1).astype(np.float64)
        four, ps = o.log_loss(x)
        z = synth_inputs(f'noise:{H}x{Wd}x{D}', B, seed=2).astype(np.float64)
        xs = o.call(z, -1)
        out = {"config": np.array(json.dumps(cfg)), "xy": x, "zy": ps['zy'], "logdet": ps['logdet'],
               "ll_z": ps['ll_z'], "ll_y": ps['ll_y'], "loss4": np.array(four), "z_in": z, "x_sampled": xs}
        for i, w in enumerate(W):
            for net in ("A", "b"):
                for k, v in w[net].items():
                    out[f"w.{i}.{net}.{k}"] = np.asarray(v, np.float32)
        np.savez_compressed(os.path.join(OUT, f"flow_{name}.npz"), **out)
    # BASELINE config 2 at its full batch (256): a digest instead of the full tensors (per-sample scalars, every 7th
    # element of zy / of the samples, per-sample fp64 sums).  Inputs and weights are regenerated from their seeds.
    cfg2 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
                num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
    o = FlowOracle(**cfg2, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=0))
    B = 256
    x = synth_inputs('cfg2', B, seed=0).astype(np.float64)
    four, ps = o.log_loss(x)
    z = synth_inputs('noise:28x28x2', B, seed=1).astype(np.float64)
    z[..., 1:] = x[..., 1:]
    xs = o.call(z, -1)
    flat = lambda a: a.reshape(B, -1)
    np.savez_compressed(os.path.join(OUT, "cfg2_b256_digest.npz"), config=np.array(json.dumps(cfg2)), weights_seed=0,
                        xy_seed=0, z_seed=1, stride=7, loss4=np.array(four), logdet=ps['logdet'], ll_z=ps['ll_z'],
                        ll_y=ps['ll_y'], zy_strided=flat(ps['zy'])[:, ::7].astype(np.float32),
                        zy_sum=flat(ps['zy']).sum(1), zy_abs=np.abs(flat(ps['zy'])).sum(1),
                        xs_strided=flat(xs)[:, ::7].astype(np.float32), xs_sum=flat(xs).sum(1),
                        xs_abs=np.abs(flat(xs)).sum(1))
    # toy
    n, width, nl = 12, 16, 2
    W = toy_init_weights(n, width, nl, seed=5, scale=1.0)
    order = list(np.random.default_rng(2).permutation(n))
    o = ToyOracle(3, 2, n, W, mask_indices=order, dtype=np.float64)
    rng = np.random.default_rng(3)
    xy = np.concatenate([rng.standard_normal((64, 2)), np.where(rng.uniform(size=(64, 1)) < 0.5, -1.0, 1.0)], 1)
    four, ps = o.log_loss(xy)
    inv, _ = o.call(xy, 1)
    out = {"n": n, "width": width, "num_layers": nl, "order": np.array(order), "xy": xy, "zy": ps['zy'],
           "logdet": ps['logdet'], "loss4": np.array(four), "inv": inv}
    for j in range(n):
        for net in ("A", "b"):
            for i, (Wm, bv) in enumerate(W[j][net]):
                out[f"w.{j}.{net}.{i}.W"] = Wm
                out[f"w.{j}.{net}.{i}.b"] = bv
    np.savez_compressed(os.path.join(OUT, "toy.npz"), **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
