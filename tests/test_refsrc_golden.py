"""Parity against the reference's OWN source files.  tests/golden/refsrc_*.npz hold what /root/reference/conv_cINN_make_model.py
and conv_cINN_base_functions.py -- imported unmodified and executed in fp64 under the NumPy TensorFlow stand-in of
oracle/tf_shim -- compute for seeded weights and inputs (oracle/make_ref_golden.py made them; README in oracle/tf_shim says what
that does and does not pin).  Checked here:

* the oracle restatement equals the reference's code to 1e-12 (fp64) on zy, the four loss scalars, the batch-mean log-det and
  the samples -- under BOTH executions of the `Lambda` closure of F:402 (`trace_once`: every group reads its own slice, the
  reading the CUDA kernels implement; `replay`: tf.keras >= 2.4 re-runs the lambda with the loop variable's final value);
* the product's Keras name map (keras_interchange.keras_weight_names) lists the variables in the order the reference's code
  creates them;
* (GPU) the CUDA path with the fixture's Keras-named weights loaded through import_keras_npz reproduces the `trace_once`
  outputs within 1e-4 (fp32, the tolerance of north_star).
"""
import glob
import json
import os

import numpy as np
import pytest
import torch

import oracle.flow_torch as FT
from oracle.flow_torch import FlowOracle

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refsrc_*.npz"))
                if os.path.basename(p) not in ("refsrc_toy.npz", "refsrc_data.npz", "refsrc_masks.npz"))   # conv-model cases
# the constructor's optional switches (LAYER_NORM = False, DILATIONS = False, lambda_y) are oracle-level fixtures
GOLDEN_GPU = [p for p in GOLDEN if not os.path.basename(p).startswith("refsrc_opt_")]
TOL64 = 1e-12


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def load(path):
    z = np.load(path)
    cfg = json.loads(str(z['cfg']))
    names = [str(n) for n in z['weight_names']]
    return z, cfg, names


def oracle_weights(cfg, z, names):
    """the fixture's Keras-named arrays in the oracle's structure, through the product's name map"""
    from arl_conditional_normalizing_flows_b200.keras_interchange import keras_weight_names
    from oracle.weights import init_weights
    o = FlowOracle(**cfg, dtype=torch.float64)
    W = init_weights(o.plan, 'init', seed=0, ln=cfg.get('LAYER_NORM', True))   # structure only; every entry is overwritten
    table = keras_weight_names([{n: list(w) for n, w in lw.items()} for lw in W])
    assert [k for k, *_ in table] == names, "Keras creation order of the reference differs from the product's name map"
    for k, ci, net, our in table:
        a = np.asarray(z[k])
        assert a.shape == np.shape(W[ci][net][our]), (k, a.shape)
        W[ci][net][our] = a
    o.set_weights(W)
    return o


def test_fixtures_are_present():
    assert len(GOLDEN) >= 3, "tests/golden/refsrc_*.npz are committed fixtures (oracle/make_ref_golden.py)"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
@pytest.mark.parametrize("mode", ["trace_once", "replay"])
def test_oracle_equals_the_reference_source(path, mode):
    z, cfg, names = load(path)
    o = oracle_weights(cfg, z, names)
    xy, zy_in = z['xy'].astype(np.float64), z['zy_in'].astype(np.float64)
    FT.LAMBDA_LATE_BINDING = mode == "replay"
    try:
        four, ps = o.log_loss(xy)
        xs = o.call(zy_in, -1)
        back = o.call(ps['zy'], -1)
    finally:
        FT.LAMBDA_LATE_BINDING = False
    assert rel(ps['zy'], z[f'{mode}:zy']) <= TOL64
    assert rel(four, z[f'{mode}:loss4']) <= TOL64
    assert rel(xs, z[f'{mode}:sample']) <= TOL64
    assert abs(float(np.mean(ps['logdet'])) - float(z[f'{mode}:logdet_mean'])) <= TOL64 * abs(float(z[f'{mode}:logdet_mean']))
    # per-sample log-dets: the reference returns only their batch mean (quirk Q1), so the fixture holds batch-of-one calls
    assert rel(ps['logdet'], z[f'{mode}:logdet_ps']) <= TOL64
    assert np.abs(back - xy).max() <= 1e-9            # and the restated inverse undoes the restated forward


def test_the_two_readings_of_the_lambda_closure_differ():
    """the question is not academic: with more than one group the two executions give different flows"""
    for path in GOLDEN:
        z, _, _ = load(path)
        assert rel(z['replay:zy'], z['trace_once:zy']) > 1e-2


@pytest.mark.skipif(not os.path.isdir(os.environ.get("CNF_REFERENCE_DIR", "/root/reference")),
                    reason="the reference checkout only exists in the build container")
def test_fixture_is_reproducible_from_the_reference_checkout(tmp_path):
    """re-run the reference's source under the shim (smallest case, into a scratch directory) and compare with the committed
    fixture"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import oracle.make_ref_golden as G; "
            "M, KL = G.load_reference(); G.run_case('card4', G.CASES['card4'], M, KL, seed=12, out_dir=%r)"
            % (root, str(tmp_path)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    z_new = np.load(os.path.join(str(tmp_path), "refsrc_card4.npz"))
    z_old = np.load(os.path.join(root, "tests", "golden", "refsrc_card4.npz"))
    assert sorted(z_new.files) == sorted(z_old.files)
    for k in z_old.files:
        if z_old[k].dtype.kind in "fc":
            np.testing.assert_allclose(z_new[k], z_old[k], rtol=1e-12, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN_GPU, ids=[os.path.basename(p)[7:-4] for p in GOLDEN_GPU])
def test_cuda_path_equals_the_reference_source(path):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    from arl_conditional_normalizing_flows_b200.keras_interchange import import_keras_npz
    z, cfg, names = load(path)
    dev = torch.device("cuda:0")
    m = cFlow(**cfg, device=dev)
    assert import_keras_npz(m, path) == len(names)
    xy = torch.from_numpy(z['xy']).to(dev)
    zy_in = torch.from_numpy(z['zy_in']).to(dev)
    four = [float(t) for t in m.log_loss(xy)]
    zy = m.last_per_sample['zy'].cpu().numpy()
    ld = float(m.last_per_sample['logdet'].double().mean())
    xs = m(zy_in, -1).cpu().numpy()
    np.testing.assert_allclose(four, z['trace_once:loss4'], rtol=1e-4)
    assert rel(zy, z['trace_once:zy']) <= 1e-4
    assert rel(xs, z['trace_once:sample']) <= 1e-4
    assert abs(ld - float(z['trace_once:logdet_mean'])) <= 1e-4 * abs(float(z['trace_once:logdet_mean']))


@pytest.mark.skipif(not os.path.isdir(os.environ.get("CNF_REFERENCE_DIR", "/root/reference")),
                    reason="the reference checkout only exists in the build container")
def test_lambda_probe_tells_the_two_executions_apart():
    """tools/tf_dump_reference.py probe (meant for a TensorFlow box) run against the stand-in's two settings"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "tf_dump_reference.py"), "probe", "--impl", "shim"],
                         capture_output=True, text=True, timeout=120, cwd=root, env=dict(os.environ, PYTHONPATH=root))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("shim")]
    assert lines[0].endswith("trace_once") and lines[1].endswith("replay"), out.stdout


def test_toy_oracle_equals_the_reference_source():
    """TOYcINN_make_model.cINN_affine (T:105-506) run under the stand-in (tests/golden/refsrc_toy.npz) against oracle/toy.py:
    12 coupling layers in an explicit mask order, call(-1), call(+1) and log_loss; Dense variables in Keras creation order
    (per coupling layer the b-net's layers, then the A-net's)."""
    from oracle.toy import ToyOracle, toy_init_weights
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refsrc_toy.npz"))
    order = [int(v) for v in z['mask_indices']]
    n, width, depth, seed = len(order), int(z['width']), int(z['depth']), int(z['seed'])
    W = toy_init_weights(n, width, depth, seed=seed, scale=1.0)
    names = [str(v) for v in z['weight_names']]
    per_layer = 2 * 2 * (depth + 2)                      # two nets x (kernel, bias) x Dense layers
    assert len(names) == n * per_layer and names[0] == "dense/kernel:0" and names[1] == "dense/bias:0"
    o = ToyOracle(3, 2, n, W, mask_indices=order, dtype=np.float64)
    xy, zy_in = z['xy'].astype(np.float64), z['zy_in'].astype(np.float64)
    zy, ld = o.call(xy, -1)
    four, _ = o.log_loss(xy)
    xs, _ = o.call(zy_in, 1)
    assert rel(zy, z['zy']) <= TOL64 and rel(ld, z['logdet']) <= TOL64
    assert rel(four, z['loss4']) <= TOL64 and rel(xs, z['sample']) <= TOL64
    assert np.abs(zy - xy).max() > 0.1                    # the flow does something


def test_data_helpers_equal_the_reference_source():
    """down / up (F:74-164, batched and single image, odd height cropped), preprocess_dataset_class with LOGITS (F:174-231),
    preprocess_dataset_SR for both model types with and without RESIDUAL (F:233-279) and de_logitify (F:287-318), as the
    reference's own functions compute them under the stand-in (tests/golden/refsrc_data.npz), against oracle/data_np.py --
    which in turn is what the CUDA data kernels are held to bit for bit (tests/test_gpu_data.py)."""
    from oracle import data_np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refsrc_data.npz"))
    img, hr = z['img'], z['hr']
    assert rel(data_np.down(img), z['down']) <= TOL64 and rel(data_np.down(img[0]), z['down1']) <= TOL64
    assert data_np.down(img).shape == (3, 4, 5, 2)
    assert rel(data_np.up(img), z['up']) <= TOL64 and rel(data_np.up(img[0]), z['up1']) <= TOL64
    assert rel(data_np.preprocess_class_logits(img, 0.01), z['logits']) <= TOL64
    assert rel(data_np.de_logitify(z['logits'], 0.01), z['delogit']) <= TOL64
    assert rel(z['delogit'], img) <= 1e-9                                  # and the reference's pair is an inverse pair
    for mt in ('SR4,2', 'SR2,1'):
        for res in (True, False):
            assert rel(data_np.preprocess_SR(hr, mt, RESIDUAL=res), z[f'sr:{mt}:{int(res)}']) <= TOL64


def test_masks_equal_the_reference_source_bit_for_bit():
    """coupling_layer.mask / decompress_mask of the reference (M:500-1073; all four masks, compressed and uncompressed, even
    and odd depths) as executed under the stand-in (tests/golden/refsrc_masks.npz) against oracle/masks_np.py, to which the CUDA
    mask kernels and the folded addressing are held bit for bit (tests/test_gpu_parity.py)."""
    from oracle import masks_np
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refsrc_masks.npz"))
    shapes = [k[3:] for k in z.files if k.startswith("uv:")]
    assert len(shapes) == 4
    for shp in shapes:
        uv = z[f"uv:{shp}"]
        for m in range(4):
            assert np.array_equal(masks_np.mask(uv, m, False), z[f"full:{shp}:{m}"])
            comp = masks_np.mask(uv, m, True)
            assert np.array_equal(comp, z[f"comp:{shp}:{m}"])
            assert np.array_equal(masks_np.decompress_mask(comp, m, uv.shape), z[f"back:{shp}:{m}"])


def _planner_cases():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refsrc_planner.json")
    return json.load(open(p)) if os.path.exists(p) else {}


@pytest.mark.parametrize("name", sorted(_planner_cases()))
def test_planners_equal_the_reference_constructor(name):
    """cFlow.__init__ of the reference (M:1431-1695) run under the stand-in for BASELINE configs 2-5 and a two-squeeze case
    (tests/golden/refsrc_planner.json): scale / factor bookkeeping, per-block shapes, auto-derived dilations, layer order,
    per-layer mask / kernels / cardinality / compressed shape and the number of Keras variables -- against BOTH the oracle
    planner and the product's C planner (csrc/plan.cpp through cFlow on the CPU)."""
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    from arl_conditional_normalizing_flows_b200.keras_interchange import keras_weight_names
    from oracle.planner import plan_flow
    rec = _planner_cases()[name]
    cfg = rec['cfg']
    p = plan_flow(cfg['io_shape'], cfg['x_d'], cfg['squeeze_factor_block_list'], cfg['ResNeXt_block_list'],
                  cfg['num_kernels_list'], cfg['cardinality_list'])
    m = cFlow(**cfg, device="cpu")
    for got_scale, got_npf, got_shapes, got_dil in (
            (p['scale_list'], p['num_prev_factors_list'], p['io_shape_list'], p['dilations_list']),
            (list(m.scale_list), list(m.num_prev_factors_list), m.io_shape_list.tolist(), m.dilations_list)):
        assert [int(v) for v in got_scale] == rec['scale_list']
        assert [int(v) for v in got_npf] == rec['num_prev_factors_list']
        assert [[int(v) for v in r] for r in got_shapes] == rec['io_shape_list']
        assert [{k: [int(d) for d in v] for k, v in d.items()} for d in got_dil] == rec['dilations_list']
    assert [L['type'] for L in p['layers']] == [L['type'] for L in rec['layers']]
    assert len(m.squeeze_factor_layers_list) == rec['n_squeeze_factor_layers']
    ref_c = [L for L in rec['layers'] if L['type'] == 'coupling']
    ora_c = [L for L in p['layers'] if L['type'] == 'coupling']
    assert len(ref_c) == len(ora_c) == len(m.coupling_layers)
    for R, L, layer in zip(ref_c, ora_c, m.coupling_layers):
        i = layer._info
        assert (R['mask'], R['mask_complement'], R['nk'], R['cardinality'], R['R']) == \
               (L['mask'], L['mask_complement'], L['nk'], L['cardinality'], L['R']) == \
               (i.mask, i.mask_complement, i.nk, i.cardinality, i.R)
        assert R['dilations'] == [int(d) for d in L['dilations']] == [i.dilation[j] for j in range(i.n_branches)]
        assert R['compressed'] == [L['h'], L['w'], L['c1']] == [i.h, i.w, i.c1]
        assert R['out_A'] == [L['h'], L['w'], L['c2']] == [i.h, i.w, i.c2]
    for L in rec['layers']:
        if L['type'] == 'factor':
            assert L['num_prev_factors'] in rec['num_prev_factors_list']
    W = m.get_weights()
    assert len(keras_weight_names([{n: list(w) for n, w in lw.items()} for lw in W])) == rec['n_variables']
