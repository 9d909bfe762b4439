"""Golden vectors from the reference's OWN source files, executed under the NumPy TensorFlow stand-in of oracle/tf_shim
(README there): /root/reference/conv_cINN_make_model.py and conv_cINN_base_functions.py are imported unmodified, the reference
`cFlow` is built, seeded weights are assigned to its Keras variables BY NAME through the product's name map
(keras_interchange.keras_weight_names: this is also the test of that map against the reference's real construction order), and
`call(+1)`, `call(-1)`, `log_loss` run eagerly in fp64.

Writes tests/golden/refsrc_<case>.npz with: the config, the Keras-named fp32 weights, the inputs, and the outputs under both
executions of the `Lambda` closure of F:402 (`*_trace_once`: every group reads its own slice; `*_replay`: late-bound `j`).
Needs /root/reference, so it runs in the build container only; the fixtures travel.  TEST INFRASTRUCTURE.

usage: python oracle/make_ref_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CNF_REFERENCE_DIR", "/root/reference")

CASES = {
    # one squeeze, two residual blocks in the first coupling block, odd channel count (D = 3, x_d = 2)
    'small_sq': dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
                     num_kernels_list=[16, 8], cardinality_list=[2, 2]),
    # the block pattern of BASELINE config 2 ([0, 1, 0, 0]) at 8x8x2 with narrow nets
    'cfg2_pattern': dict(io_shape=[8, 8, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[1, 1, 1, 1],
                         num_kernels_list=[16, 16, 8, 8], cardinality_list=[2, 2, 2, 2]),
    # four groups: the Lambda question matters most here
    'card4': dict(io_shape=[4, 4, 2], x_d=1, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
                  num_kernels_list=[16], cardinality_list=[4]),
    # the constructor's optional switches (oracle-only fixtures: names start with "opt_")
    'opt_no_ln': dict(io_shape=[8, 8, 2], x_d=1, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[1, 2],
                      num_kernels_list=[16, 8], cardinality_list=[2, 2], LAYER_NORM=False),
    # (DILATIONS=False cannot be a case: the reference's constructor reads self.dilations_list, which it only defines when
    #  DILATIONS is true, and raises AttributeError at M:1643 -- the product mirrors that)
    'opt_lambda_y': dict(io_shape=[4, 4, 4], x_d=2, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
                         num_kernels_list=[8], cardinality_list=[2], lambda_y=7),
}


def load_reference():
    """import the reference model module against the shim (np.int: NumPy < 1.24 alias the reference still uses, M:1532)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
    sys.path.insert(0, REF)
    if not hasattr(np, "int"):
        np.int = int
    import conv_cINN_make_model as M          # noqa: E402  (the reference's file)
    from tensorflow.keras import layers as KL  # noqa: E402  (the shim)
    return M, KL


def run_case(name, cfg, M, KL, seed, out_dir=None):
    sys.path.insert(0, ROOT)
    from arl_conditional_normalizing_flows_b200.keras_interchange import keras_weight_names
    from oracle.flow_torch import FlowOracle
    from oracle.weights import init_weights
    import torch

    KL.reset_state()                       # a fresh process as far as Keras' name counters are concerned
    ref = M.cFlow(**cfg)
    variables = {v.name: v for v in KL.all_variables()}

    oracle = FlowOracle(**cfg, dtype=torch.float64)
    W = init_weights(oracle.plan, 'rand', seed=seed, ln=cfg.get('LAYER_NORM', True))
    table = keras_weight_names([{n: list(w) for n, w in lw.items()} for lw in W])
    names = [k for k, *_ in table]
    assert len(set(names)) == len(names)
    assert set(names) == set(variables), (sorted(set(names) ^ set(variables))[:6], len(names), len(variables))
    # Keras creation order == the order the name map walks
    assert names == [v.name for v in KL.all_variables()], "creation order of the reference differs from the name map"
    keras_w = {}
    for k, ci, net, our in table:
        a = np.asarray(W[ci][net][our], dtype=np.float32)
        assert tuple(variables[k].shape) == a.shape, (k, variables[k].shape, a.shape)
        variables[k].assign(a)
        keras_w[k] = a

    rng = np.random.default_rng(seed + 1)
    H, Wd, D = cfg['io_shape']
    xd = cfg['x_d']
    B = 3
    xy = rng.standard_normal((B, H, Wd, D)).astype(np.float32).astype(np.float64)
    zy_in = rng.standard_normal((B, H, Wd, D)).astype(np.float32).astype(np.float64)
    zy_in[..., xd:] = xy[..., xd:]

    out = {}
    for mode, replay in (('trace_once', False), ('replay', True)):
        KL.LAMBDA_REPLAY = replay
        zy, ld = ref(xy, 1)
        four = [float(v) for v in ref.log_loss(xy)]
        xs = ref(zy_in, -1)
        back = ref(zy, -1)
        # the reference only returns the batch MEAN of the log-dets (M:1325, quirk Q1): per-sample values from batch-of-one calls
        ld_ps = [float(ref(xy[i:i + 1], 1)[1]) for i in range(B)]
        out[mode] = dict(zy=np.asarray(zy), logdet_mean=float(ld), logdet_ps=np.asarray(ld_ps), loss4=np.asarray(four),
                         sample=np.asarray(xs), roundtrip=float(np.abs(np.asarray(back) - xy).max()))

    # the oracle on the same weights and inputs, under both readings of the Lambda closure
    import oracle.flow_torch as FT
    oracle.set_weights(W)
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max()))  # noqa: E731
    report = {}
    for mode, late in (('trace_once', False), ('replay', True)):
        FT.LAMBDA_LATE_BINDING = late
        four_o, ps = oracle.log_loss(xy)
        xs_o = oracle.call(zy_in, -1)
        t = out[mode]
        report.update({f"{mode}:zy": rel(ps['zy'], t['zy']), f"{mode}:loss4": rel(four_o, t['loss4']),
                       f"{mode}:sample": rel(xs_o, t['sample']),
                       f"{mode}:logdet_mean": abs(float(np.mean(ps['logdet'])) - t['logdet_mean']) / abs(t['logdet_mean']),
                       f"{mode}:logdet_ps": rel(ps['logdet'], t['logdet_ps'])})
    FT.LAMBDA_LATE_BINDING = False
    t = out['trace_once']
    report['replay_vs_trace_once'] = rel(out['replay']['zy'], t['zy'])
    print(f"{name}: {len(names)} Keras variables; oracle vs reference source: " +
          ", ".join(f"{k} {v:.2e}" for k, v in report.items()) +
          f"; round trips {t['roundtrip']:.1e} / {out['replay']['roundtrip']:.1e}")
    path = os.path.join(out_dir or os.path.join(ROOT, "tests", "golden"), f"refsrc_{name}.npz")
    np.savez_compressed(path, cfg=json.dumps(cfg), xy=xy.astype(np.float32), zy_in=zy_in.astype(np.float32),
                        weight_names=np.array(names), **keras_w,
                        **{f"{mode}:{k}": np.asarray(v) for mode, d in out.items() for k, v in d.items()})
    return report


def run_toy(KL, seed=20):
    """TOYcINN_make_model.cINN_affine (T:105-506) under the shim against oracle/toy.py: 12 coupling layers with an explicit
    mask order, Dense weights assigned in Keras creation order (per layer: the b-net's Dense layers, then the A-net's)."""
    import TOYcINN_make_model as T            # the reference's file
    sys.path.insert(0, ROOT)
    from oracle.toy import ToyOracle, toy_init_weights
    KL.reset_state()
    n, width, depth = 12, 16, 2
    order = [int(v) for v in np.random.default_rng(seed).permutation(n)]
    ref = T.cINN_affine(3, 2, n, width, depth, None, mask_indices=order)
    W = toy_init_weights(n, width, depth, seed=seed, scale=1.0)
    variables = KL.all_variables()
    flat, k = {}, 0
    for j in range(n):
        for net in ('b', 'A'):                 # T:52-66 builds the b-net first
            for Wm, bv in W[j][net]:
                for a in (Wm, bv):
                    v = variables[k]
                    assert tuple(v.shape) == a.shape, (v.name, v.shape, a.shape)
                    v.assign(a)
                    flat[v.name] = np.asarray(a, np.float32)
                    k += 1
    assert k == len(variables)
    rng = np.random.default_rng(seed + 1)
    B = 64
    xy = np.concatenate([rng.standard_normal((B, 2)), np.where(rng.uniform(size=(B, 1)) < 0.5, -1.0, 1.0)], 1)
    xy = xy.astype(np.float32).astype(np.float64)
    zy_in = rng.standard_normal((B, 3)).astype(np.float32).astype(np.float64)
    zy, ld = ref(xy, -1)
    four = [float(v) for v in ref.log_loss(xy)]
    xs, _ = ref(zy_in, 1)
    o = ToyOracle(3, 2, n, W, mask_indices=order, dtype=np.float64)
    zy_o, ld_o = o.call(xy, -1)
    four_o, _ = o.log_loss(xy)
    xs_o, _ = o.call(zy_in, 1)
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max()))  # noqa: E731
    print(f"toy: {len(variables)} Keras variables; oracle vs reference source: zy {rel(zy_o, zy):.2e}, logdet {rel(ld_o, ld):.2e}, "
          f"loss4 {rel(four_o, four):.2e}, sample {rel(xs_o, xs):.2e}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refsrc_toy.npz"), mask_indices=np.array(order), width=width,
                        depth=depth, seed=seed, xy=xy.astype(np.float32), zy_in=zy_in.astype(np.float32), zy=np.asarray(zy),
                        logdet=np.asarray(ld), loss4=np.asarray(four), sample=np.asarray(xs),
                        weight_names=np.array([v.name for v in variables]))


def run_data():
    """the data-side helpers of conv_cINN_base_functions.py (down / up F:74-164, preprocess_dataset_class F:174-231,
    preprocess_dataset_SR F:233-279, de_logitify F:287-318) under the shim against oracle/data_np.py"""
    import tensorflow as tf
    import conv_cINN_base_functions as F      # the reference's file
    sys.path.insert(0, ROOT)
    from oracle import data_np
    rng = np.random.default_rng(30)
    img = rng.uniform(0, 1, (3, 9, 10, 2)).astype(np.float32).astype(np.float64)      # odd height: the crop of F:94-97
    hr = rng.uniform(0, 1, (2, 16, 16, 3)).astype(np.float32).astype(np.float64)
    out = dict(img=img, hr=hr)
    out['down'] = F.down(img)
    out['down1'] = F.down(img[0])
    out['up'] = F.up(img)
    out['up1'] = F.up(img[0])
    out['logits'] = np.stack(list(F.preprocess_dataset_class(tf.data.Dataset.from_tensor_slices(img), LOGITS=True, a=0.01)))
    for mt in ('SR4,2', 'SR2,1'):
        for res in (True, False):
            out[f'sr:{mt}:{int(res)}'] = np.stack(list(F.preprocess_dataset_SR(tf.data.Dataset.from_tensor_slices(hr), mt, RESIDUAL=res)))
    out['delogit'] = F.de_logitify(np.array(out['logits']), a=0.01)
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(1e-300, np.abs(np.asarray(b)).max()))  # noqa: E731
    errs = dict(down=rel(data_np.down(img), out['down']), down1=rel(data_np.down(img[0]), out['down1']),
                up=rel(data_np.up(img), out['up']), up1=rel(data_np.up(img[0]), out['up1']),
                logits=rel(data_np.preprocess_class_logits(img, 0.01), out['logits']),
                delogit=rel(data_np.de_logitify(out['logits'], 0.01), out['delogit']))
    for mt in ('SR4,2', 'SR2,1'):
        for res in (True, False):
            errs[f'sr:{mt}:{int(res)}'] = rel(data_np.preprocess_SR(hr, mt, RESIDUAL=res), out[f'sr:{mt}:{int(res)}'])
    print("data helpers: oracle vs reference source: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refsrc_data.npz"), **{k: np.asarray(v) for k, v in out.items()})


PLANNER_CASES = {
    'cfg2': dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
                 num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4]),
    'cfg3': dict(io_shape=[32, 32, 4], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
                 num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4]),
    'cfg4': dict(io_shape=[64, 64, 6], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[1] * 4,
                 num_kernels_list=[64, 64, 32, 32], cardinality_list=[4, 4, 2, 2]),
    'cfg5': dict(io_shape=[128, 128, 4], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[1] * 4,
                 num_kernels_list=[64, 64, 32, 32], cardinality_list=[2, 2, 2, 2]),
    'sq2': dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 1], ResNeXt_block_list=[2, 1],
                num_kernels_list=[16, 8], cardinality_list=[2, 2]),
}


def run_planner(M, KL):
    """cFlow.__init__ (M:1431-1695) of the reference for the BASELINE configs (R = 1 on the two large ones: the planner does
    not depend on it): every derived attribute the constructor computes, next to oracle/planner.py"""
    sys.path.insert(0, ROOT)
    from oracle.planner import plan_flow
    out = {}
    for name, cfg in PLANNER_CASES.items():
        KL.reset_state()
        ref = M.cFlow(**cfg)
        layers = []
        for L in ref.layers_list:
            kind = type(L).__name__
            if kind == 'coupling_layer':
                layers.append(dict(type='coupling', mask=int(L.which_mask), mask_complement=int(L.which_mask_complement),
                                   nk=int(L.num_kernels), cardinality=int(L.cardinality), R=int(L.num_res_blocks),
                                   dilations=[int(d) for d in L.which_dilations],
                                   in_shape=[int(L.input_height), int(L.input_width), int(L.input_depth)],
                                   compressed=[int(L.compressed_height), int(L.compressed_width), int(L.compressed_depth)],
                                   n_variables_A=len([w for lay in _layers_of(L.model_A) for w in lay.weights]),
                                   out_A=[int(v) for v in L.model_A.outputs.shape[1:]]))
            elif kind == 'squeeze_layer':
                layers.append(dict(type='squeeze'))
            else:
                layers.append(dict(type='factor', num_prev_factors=int(L.num_prev_factors)))
        rec = dict(cfg=cfg, scale_list=[int(v) for v in ref.scale_list],
                   num_prev_factors_list=[int(v) for v in ref.num_prev_factors_list],
                   io_shape_list=[[int(v) for v in r] for r in ref.io_shape_list],
                   dilations_list=[{k: [int(d) for d in v] for k, v in d.items()} for d in ref.dilations_list],
                   layers=layers, n_squeeze_factor_layers=len(ref.squeeze_factor_layers_list),
                   n_variables=len(KL.all_variables()))
        p = plan_flow(cfg['io_shape'], cfg['x_d'], cfg['squeeze_factor_block_list'], cfg['ResNeXt_block_list'],
                      cfg['num_kernels_list'], cfg['cardinality_list'])
        assert rec['scale_list'] == p['scale_list'] and rec['num_prev_factors_list'] == p['num_prev_factors_list'], name
        assert rec['io_shape_list'] == p['io_shape_list'], name
        assert [L['type'] for L in layers] == [L['type'] for L in p['layers']], name
        out[name] = rec
        print(f"planner {name}: {len(layers)} layers, {rec['n_variables']} Keras variables, scales {rec['scale_list']}, "
              f"dilations {[d['channelwise'] for d in rec['dilations_list']]}")
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "refsrc_planner.json"), "w"), indent=1)


def _layers_of(model):
    """the layers of a functional shim model, each once"""
    seen, order = set(), []

    def walk(t):
        if t.node is None:
            return
        layer, ins = t.node
        for v in (ins if isinstance(ins, (list, tuple)) else [ins]):
            walk(v)
        if id(layer) not in seen:
            seen.add(id(layer))
            order.append(layer)
    for t in (model.outputs if isinstance(model.outputs, (list, tuple)) else [model.outputs]):
        walk(t)
    return order


def run_masks(M, KL):
    """coupling_layer.mask / decompress_mask (M:500-1073) of the reference, all four masks, compressed and not, even and odd
    depths, against the literal transcription in oracle/masks_np.py (bit exact: pure index permutations)"""
    sys.path.insert(0, ROOT)
    from oracle import masks_np
    import tensorflow as tf
    rng = np.random.default_rng(40)
    out, worst = {}, 0.0
    for (H, W, D) in ((4, 4, 2), (4, 6, 3), (6, 4, 4), (2, 2, 5)):
        uv = rng.standard_normal((2, H, W, D)).astype(np.float32).astype(np.float64)
        out[f'uv:{H}x{W}x{D}'] = uv
        for m in range(4):
            KL.reset_state()
            L = M.coupling_layer([H, W, D], m, 1, 2, 8, 3, tf.keras.initializers.Orthogonal(gain=0.1), LAYER_NORM=True,
                                 which_dilations=[1])
            full = np.asarray(L.mask(uv, m, False))
            comp = np.asarray(L.mask(uv, m, True))
            back = np.asarray(L.decompress_mask(comp, m, (2, H, W, D)))
            out[f'full:{H}x{W}x{D}:{m}'], out[f'comp:{H}x{W}x{D}:{m}'], out[f'back:{H}x{W}x{D}:{m}'] = full, comp, back
            assert np.array_equal(masks_np.mask(uv, m, False), full), (H, W, D, m)
            assert np.array_equal(masks_np.mask(uv, m, True), comp), (H, W, D, m)
            assert np.array_equal(masks_np.decompress_mask(comp, m, (2, H, W, D)), back), (H, W, D, m)
            assert np.array_equal(back, full)                      # decompress(compress) = the masked tensor
            worst = max(worst, float(np.abs(back - full).max()))
    print(f"masks: {len(out)} arrays, oracle/masks_np.py bit-identical to the reference's mask / decompress_mask")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "refsrc_masks.npz"), **out)


if __name__ == "__main__":
    M, KL = load_reference()
    run_masks(M, KL)
    run_planner(M, KL)
    run_data()
    for i, (name, cfg) in enumerate(CASES.items()):
        run_case(name, cfg, M, KL, seed=10 + i)
    run_toy(KL)
