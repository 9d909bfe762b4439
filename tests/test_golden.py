"""Committed golden vectors (tests/golden, made by oracle/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA path must match them on the GPU box (which has no /root/reference)."""
import numpy as np
import pytest
import torch

from golden_util import load_flow, load_toy
from oracle.flow_torch import FlowOracle
from oracle.toy import ToyOracle


@pytest.mark.parametrize("name", ["tiny", "small_sq"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 1e-4)])
def test_oracle_reproduces_golden(name, dtype, tol):
    cfg, W, d = load_flow(name)
    o = FlowOracle(**cfg, weights=W, dtype=dtype)
    four, ps = o.log_loss(d["xy"])
    s = lambda a: np.abs(a).max()
    assert np.abs(ps['zy'] - d["zy"]).max() <= tol * s(d["zy"])
    np.testing.assert_allclose(ps['logdet'], d["logdet"], rtol=tol, atol=tol * s(d["logdet"]))
    np.testing.assert_allclose(ps['ll_z'], d["ll_z"], rtol=tol)
    np.testing.assert_allclose(four, d["loss4"], rtol=tol, atol=tol * s(d["logdet"]))
    xs = o.call(d["z_in"], -1)
    assert np.abs(xs - d["x_sampled"]).max() <= tol * s(d["x_sampled"])


def test_toy_oracle_reproduces_golden():
    W, d = load_toy()
    o = ToyOracle(3, 2, int(d["n"]), W, mask_indices=list(d["order"]), dtype=np.float64)
    four, ps = o.log_loss(d["xy"])
    np.testing.assert_allclose(ps['zy'], d["zy"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(four, d["loss4"], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny", "small_sq"])
def test_cuda_matches_golden(name):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    cfg, W, d = load_flow(name)
    m = cFlow(**cfg, device="cuda:0")
    m.set_weights(W)
    x = torch.from_numpy(d["xy"].astype(np.float32)).cuda()
    four = [float(t) for t in m.log_loss(x)]
    tol = 1e-4
    s = lambda a: np.abs(a).max()
    zy = m.last_per_sample['zy'].cpu().numpy()
    assert np.abs(zy - d["zy"]).max() <= tol * s(d["zy"])
    np.testing.assert_allclose(m.last_per_sample['logdet'].cpu().numpy(), d["logdet"], rtol=tol, atol=tol * s(d["logdet"]))
    np.testing.assert_allclose(m.last_per_sample['ll_z'].cpu().numpy(), d["ll_z"], rtol=tol)
    np.testing.assert_allclose(four, d["loss4"], rtol=tol, atol=tol * s(d["logdet"]))
    xs = m(torch.from_numpy(d["z_in"].astype(np.float32)).cuda(), -1).cpu().numpy()
    assert np.abs(xs - d["x_sampled"]).max() <= tol * s(d["x_sampled"])


@pytest.mark.gpu
def test_cuda_toy_matches_golden():
    from arl_conditional_normalizing_flows_b200.TOYcINN_make_model import cINN_affine
    W, d = load_toy()
    n = int(d["n"])
    m = cINN_affine(3, 2, n, int(d["width"]), int(d["num_layers"]), None, mask_indices=[int(i) for i in d["order"]],
                    device="cuda:0")
    for j, cl in enumerate(m.coupling_layers_list):
        flat = []
        for net in ('b', 'A'):
            for Wm, bv in W[j][net]:
                flat += [Wm, bv]
        cl.set_weights(flat)
    x = torch.from_numpy(d["xy"].astype(np.float32)).cuda()
    four = [float(t) for t in m.log_loss(x)]
    np.testing.assert_allclose(four, d["loss4"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(m.last_per_sample['zy'].cpu().numpy(), d["zy"], rtol=1e-4, atol=1e-5)
    inv, _ = m(x, 1)
    np.testing.assert_allclose(inv.cpu().numpy(), d["inv"], rtol=1e-4, atol=1e-5)


def test_cfg2_full_batch_digest_is_what_the_oracle_computes():
    """tests/golden/cfg2_b256_digest.npz (config 2, batch 256): the oracle reproduces the first samples of the digest (every
    op is per-sample, so a sub-batch gives the same per-sample numbers)."""
    import json
    import os
    import torch
    from oracle.flow_torch import FlowOracle
    from oracle.weights import init_weights, synth_inputs
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg2_b256_digest.npz"))
    cfg = json.loads(str(g['config']))
    o = FlowOracle(**cfg, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=int(g['weights_seed'])))
    n = 4
    x = synth_inputs('cfg2', 256, seed=int(g['xy_seed'])).astype(np.float64)[:n]
    _, ps = o.log_loss(x)
    np.testing.assert_allclose(ps['logdet'], g['logdet'][:n], rtol=1e-10)
    np.testing.assert_allclose(ps['ll_z'], g['ll_z'][:n], rtol=1e-10)
    np.testing.assert_allclose(ps['zy'].reshape(n, -1)[:, ::int(g['stride'])], g['zy_strided'][:n], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ps['zy'].reshape(n, -1).sum(1), g['zy_sum'][:n], rtol=1e-10)
