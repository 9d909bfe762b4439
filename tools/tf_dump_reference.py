"""Pin parity at the TensorFlow boundary: dump golden outputs of the UNMODIFIED reference (TensorFlow) for a seeded case.

Nothing in this image can run TensorFlow: the oracle in oracle/ is pinned to the reference's SOURCE (run under the
NumPy stand-in of oracle/tf_shim, tests/golden/refsrc_*) but not to TensorFlow's own kernels.  This script is the route from
there to fixtures written by a real TensorFlow.  It has three parts:

  1. anywhere this repo is checked out (no TF needed):
         python tools/tf_dump_reference.py export --config cfg2 --batch 8 --out /tmp/case_cfg2
     writes  /tmp/case_cfg2/weights_keras.npz   the seeded 'rand' weights keyed by the reference's Keras variable names
                                                (arl_conditional_normalizing_flows_b200/keras_interchange.py)
             /tmp/case_cfg2/inputs.npz          xy (synthetic batch), zy_in (fixed latent for sampling), the config

  2. on any box with TensorFlow >= 2.7, tensorflow_probability >= 0.15 and a checkout of the reference:
         python tools/tf_dump_reference.py run --reference /path/to/ARL_Conditional_Normalizing_Flows \
                --case /tmp/case_cfg2 --out tests/golden/tf_cfg2.npz
     builds conv_cINN_make_model.cFlow (M:1396-1904) with the case's hyper-parameters as the FIRST Keras model of the
     process (the name map counts layers from zero), calls it once to create the variables (C:572-576), assigns every
     variable by name, and stores  zy, log_detJ = model(xy, 1) (M:1743-1772),  the 4-tuple of model.log_loss(xy)
     (M:1800-1848),  xy_sampled = model(zy_in, -1) (M:1774-1798)  and the per-sample log-dets of the first samples
     (batch-of-one calls: the reference only returns the batch mean, Q1).

  3. on the same TensorFlow box, the one-second answer to the question oracle/tf_shim/README.md raises about F:402:
         python tools/tf_dump_reference.py probe --reference /path/to/ARL_Conditional_Normalizing_Flows
     builds the reference's grouped_convolution (two groups, all-ones kernels) as a functional Model, feeds an input whose
     FIRST group slice is ones and whose second is zeros, and prints `trace_once` (group 0 read its own slice: what the oracle
     default and the CUDA kernels implement) or `replay` (tf.keras re-ran the Lambda with the loop variable's final value:
     every group read the last slice; oracle.flow_torch.LAMBDA_LATE_BINDING = True reproduces that).
     `--impl shim` answers the same question for the two settings of the stand-in (oracle/tf_shim), as a test of the probe.

tests/test_tf_golden.py consumes every tests/golden/tf_*.npz that exists: the oracle must reproduce it to 1e-5 (that pins
the oracle) and the CUDA path to the north-star tolerance 1e-4.  `run --impl oracle` writes the same file format from the
oracle instead of TensorFlow; it exists to test the harness and must never be committed as tf_*.npz.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    'small': dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
                  num_kernels_list=[16, 8], cardinality_list=[2, 2]),
    'cfg2': dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
                 num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4]),
    'cfg3': dict(io_shape=[32, 32, 4], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
                 num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4]),
}


def _inputs(name, cfg, B, seed):
    from oracle.weights import synth_inputs
    H, W, D = cfg['io_shape']
    xy = synth_inputs(name if name in ('cfg2', 'cfg3') else f'noise:{H}x{W}x{D}', B, seed=seed)
    z = synth_inputs(f'noise:{H}x{W}x{D}', B, seed=seed + 1)
    z[..., cfg['x_d']:] = xy[..., cfg['x_d']:]
    return xy, z


def cmd_export(a):
    from oracle.planner import plan_flow
    from oracle.weights import init_weights
    from arl_conditional_normalizing_flows_b200.keras_interchange import keras_weight_names
    cfg = CONFIGS[a.config]
    plan = plan_flow(cfg['io_shape'], cfg['x_d'], cfg['squeeze_factor_block_list'], cfg['ResNeXt_block_list'],
                     cfg['num_kernels_list'], cfg['cardinality_list'])
    W = init_weights(plan, a.weights, seed=a.seed)
    table = keras_weight_names([{n: list(w) for n, w in lw.items()} for lw in W])
    os.makedirs(a.out, exist_ok=True)
    np.savez(os.path.join(a.out, 'weights_keras.npz'),
             **{k: np.asarray(W[ci][net][our], np.float32) for k, ci, net, our in table})
    xy, z = _inputs(a.config, cfg, a.batch, a.seed + 100)
    meta = dict(config=a.config, cfg=cfg, batch=a.batch, seed=a.seed, weights=a.weights)
    np.savez(os.path.join(a.out, 'inputs.npz'), xy=xy, zy_in=z, meta=json.dumps(meta))
    print(f"wrote {len(table)} variables and a batch of {a.batch} to {a.out}")


def _run_tf(a, meta, xy, z, wz):
    np.int = int                       # the reference uses np.int (M:1532), removed in NumPy 1.24
    sys.path.insert(0, a.reference)
    import tensorflow as tf
    import conv_cINN_make_model as M   # the reference's own module, unmodified
    cfg = meta['cfg']
    model = M.cFlow(io_shape=cfg['io_shape'], x_d=cfg['x_d'], squeeze_factor_block_list=cfg['squeeze_factor_block_list'],
                    ResNeXt_block_list=cfg['ResNeXt_block_list'], num_kernels_list=cfg['num_kernels_list'],
                    cardinality_list=cfg['cardinality_list'])
    model(tf.constant(xy[:1]), 1)      # builds every variable (C:572-576)
    by_name = {w.name: w for w in model.weights}
    missing = [k for k in wz.files if k not in by_name]
    extra = [k for k in by_name if k not in wz.files]
    if missing or extra:
        raise SystemExit(f"Keras variable names do not match the name map: missing {missing[:5]}, unexpected {extra[:5]} "
                         "(was another Keras model built earlier in this process?)")
    for k in wz.files:
        by_name[k].assign(wz[k].reshape(by_name[k].shape))
    zy, ld = model(tf.constant(xy), 1)
    four = model.log_loss(tf.constant(xy))
    xs = model(tf.constant(z), -1)
    ld_ps = [float(model(tf.constant(xy[i:i + 1]), 1)[1]) for i in range(min(len(xy), 4))]
    return (zy.numpy(), float(ld), [float(t) for t in four], xs.numpy(), ld_ps,
            f"tensorflow {tf.__version__}")


def cmd_probe(a):
    np.int = int
    if a.impl == 'shim':
        sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    sys.path.insert(0, a.reference)
    import tensorflow as tf
    from tensorflow.keras import Model
    from tensorflow.keras.layers import Input
    import conv_cINN_base_functions as F      # the reference's own module, unmodified
    verdicts = []
    for replay in ((False, True) if a.impl == 'shim' else (None,)):
        if replay is not None:
            from tensorflow.keras import layers as KL
            KL.LAMBDA_REPLAY = replay
        inp = Input(shape=(4, 4, 4))
        out = F.grouped_convolution(inp, 4, _strides=(1, 1), ksize=(1, 1), dilation=(1, 1), cardinality=2, init='ones')
        model = Model(inputs=inp, outputs=out)
        x = np.zeros((1, 4, 4, 4), np.float32)
        x[..., :2] = 1.0                       # only the FIRST group's input slice is non-zero
        y = np.asarray(model(tf.constant(x)))
        g0, g1 = float(y[0, 0, 0, 0]), float(y[0, 0, 0, 2])     # an output channel of group 0 / of group 1
        verdict = 'trace_once' if (g0, g1) == (2.0, 0.0) else 'replay' if (g0, g1) == (0.0, 0.0) else f'unexpected {g0} {g1}'
        who = f"tensorflow {getattr(tf, '__version__', '?')}" if replay is None else f"shim, LAMBDA_REPLAY = {replay}"
        print(f"{who}: group 0 -> {g0}, group 1 -> {g1}: {verdict}")
        verdicts.append(verdict)
    return verdicts


def _run_oracle(meta, xy, z):
    import torch
    from oracle.flow_torch import FlowOracle
    from oracle.weights import init_weights
    o = FlowOracle(**meta['cfg'], dtype=torch.float64)
    o.set_weights(init_weights(o.plan, meta['weights'], seed=meta['seed']))
    four, ps = o.log_loss(xy.astype(np.float64))
    xs = o.call(z.astype(np.float64), -1)
    return ps['zy'], float(ps['logdet'].mean()), list(four), xs, [float(v) for v in ps['logdet'][:4]], "oracle (NOT TensorFlow)"


def cmd_run(a):
    inp = np.load(os.path.join(a.case, 'inputs.npz'))
    meta = json.loads(str(inp['meta']))
    xy, z = inp['xy'], inp['zy_in']
    if a.impl == 'tf':
        wz = np.load(os.path.join(a.case, 'weights_keras.npz'))
        zy, ld, four, xs, ld_ps, how = _run_tf(a, meta, xy, z, wz)
    else:
        zy, ld, four, xs, ld_ps, how = _run_oracle(meta, xy, z)
    np.savez_compressed(a.out, zy=np.asarray(zy, np.float32), log_detJ=np.float64(ld), loss4=np.asarray(four, np.float64),
                        xy_sampled=np.asarray(xs, np.float32), logdet_first=np.asarray(ld_ps, np.float64),
                        meta=json.dumps(dict(meta, produced_by=how)))
    print(f"wrote {a.out} ({how}): loss {four[0]:.6f}, log_detJ {ld:.6f}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    sub = ap.add_subparsers(dest='cmd', required=True)
    e = sub.add_parser('export')
    e.add_argument('--config', choices=sorted(CONFIGS), default='cfg2')
    e.add_argument('--batch', type=int, default=8)
    e.add_argument('--seed', type=int, default=0)
    e.add_argument('--weights', choices=['rand', 'init'], default='rand')
    e.add_argument('--out', required=True)
    r = sub.add_parser('run')
    r.add_argument('--reference', default='/root/reference')
    r.add_argument('--case', required=True)
    r.add_argument('--out', required=True)
    r.add_argument('--impl', choices=['tf', 'oracle'], default='tf')
    q = sub.add_parser('probe')
    q.add_argument('--reference', default='/root/reference')
    q.add_argument('--impl', choices=['tf', 'shim'], default='tf')
    args = ap.parse_args()
    {'export': cmd_export, 'run': cmd_run, 'probe': cmd_probe}[args.cmd](args)
