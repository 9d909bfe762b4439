"""Consumes tests/golden/tf_*.npz: golden outputs of the UNMODIFIED TensorFlow reference written by
tools/tf_dump_reference.py on a TensorFlow box (this image has none).  While no such file is committed the
reference-boundary parity stays "unpinned" and these tests skip; the harness itself is exercised with a file the same
script produces from the oracle (never committed under the tf_ prefix)."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle.flow_torch import FlowOracle
from oracle.weights import init_weights

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "tf_*.npz")))


def _inputs(meta):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import tf_dump_reference as T
    return T._inputs(meta['config'], meta['cfg'], meta['batch'], meta['seed'] + 100)


def _check_oracle(path, tol):
    """The oracle in its default reading of the Lambda closure of F:402 (every group its own slice: what the CUDA kernels
    implement).  If a TensorFlow-written file only matches with oracle.flow_torch.LAMBDA_LATE_BINDING = True, real tf.keras
    re-runs the lambda with the loop variable's final value (oracle/tf_shim/README.md) and this assertion says so."""
    import oracle.flow_torch as FT
    g = np.load(path)
    meta = json.loads(str(g['meta']))
    xy, z = _inputs(meta)
    o = FlowOracle(**meta['cfg'], dtype=torch.float64)
    o.set_weights(init_weights(o.plan, meta['weights'], seed=meta['seed']))
    four, ps = o.log_loss(xy.astype(np.float64))
    scale = np.abs(g['zy']).max()
    if np.abs(ps['zy'] - g['zy']).max() > tol * scale:
        FT.LAMBDA_LATE_BINDING = True
        try:
            _, late = o.log_loss(xy.astype(np.float64))
        finally:
            FT.LAMBDA_LATE_BINDING = False
        if np.abs(late['zy'] - g['zy']).max() <= tol * scale:
            pytest.fail(f"{os.path.basename(path)} matches the oracle only with LAMBDA_LATE_BINDING = True: the reference's "
                        "TensorFlow executes the Lambda closure of F:402 late-bound (every group reads the last group's "
                        "channels); the default oracle and the CUDA kernels implement the per-group slices")
    assert np.abs(ps['zy'] - g['zy']).max() <= tol * scale
    np.testing.assert_allclose(ps['logdet'].mean(), float(g['log_detJ']), rtol=tol, atol=tol)
    np.testing.assert_allclose(four, g['loss4'], rtol=tol)
    np.testing.assert_allclose(ps['logdet'][:len(g['logdet_first'])], g['logdet_first'], rtol=tol, atol=tol)
    xs = o.call(z.astype(np.float64), -1)
    assert np.abs(xs - g['xy_sampled']).max() <= tol * np.abs(g['xy_sampled']).max()
    return meta


@pytest.mark.skipif(not GOLDEN, reason="no TensorFlow golden file committed yet (tools/tf_dump_reference.py): parity unpinned")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_tensorflow_golden(path):
    meta = _check_oracle(path, 2e-5)          # fp32 TensorFlow vs the fp64 restatement
    assert "tensorflow" in meta['produced_by']


@pytest.mark.gpu
@pytest.mark.skipif(not GOLDEN, reason="no TensorFlow golden file committed yet (tools/tf_dump_reference.py): parity unpinned")
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_path_reproduces_tensorflow_golden(path):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    g = np.load(path)
    meta = json.loads(str(g['meta']))
    xy, z = _inputs(meta)
    m = cFlow(**meta['cfg'], device="cuda:0")
    plan = FlowOracle(**meta['cfg'], dtype=torch.float64).plan
    m.set_weights(init_weights(plan, meta['weights'], seed=meta['seed']))
    zy, ld = m(torch.from_numpy(xy).cuda(), 1)
    assert np.abs(zy.cpu().numpy() - g['zy']).max() <= 1e-4 * np.abs(g['zy']).max()
    np.testing.assert_allclose(float(ld), float(g['log_detJ']), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose([float(t) for t in m.log_loss(torch.from_numpy(xy).cuda())], g['loss4'], rtol=1e-4)
    xs = m(torch.from_numpy(z).cuda(), -1).cpu().numpy()
    assert np.abs(xs - g['xy_sampled']).max() <= 1e-4 * np.abs(g['xy_sampled']).max()


def test_golden_harness_with_an_oracle_written_file(tmp_path):
    """export -> run --impl oracle -> the consumer above: the file format, the seeded regeneration of inputs and weights
    and the comparisons work end to end (the Keras-named weight file holds exactly the regenerated weights)."""
    case = tmp_path / "case"
    env = dict(os.environ, PYTHONPATH=ROOT)
    tool = os.path.join(ROOT, "tools", "tf_dump_reference.py")
    subprocess.run([sys.executable, tool, "export", "--config", "small", "--batch", "3", "--out", str(case)], check=True, env=env)
    out = tmp_path / "harness_small.npz"
    subprocess.run([sys.executable, tool, "run", "--impl", "oracle", "--case", str(case), "--out", str(out)], check=True, env=env)
    meta = _check_oracle(str(out), 1e-6)
    assert "NOT TensorFlow" in meta['produced_by']
    wz = np.load(case / "weights_keras.npz")
    plan = FlowOracle(**meta['cfg'], dtype=torch.float64).plan
    W = init_weights(plan, 'rand', seed=0)
    np.testing.assert_array_equal(wz['conv2d/kernel:0'], W[0]['b']['stem.kernel'])
    assert wz['tanh_scaling_layer/Variable:0'].shape == ()
