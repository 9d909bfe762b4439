"""Pins the floating-point side of the oracle (SURVEY §4 items 3-7, §8c 'what pins results')."""
import math

import numpy as np
import pytest
import torch

from oracle import nets_np
from oracle.flow_torch import FlowOracle, st_net, _t
from oracle.planner import plan_flow, plan_coupling
from oracle.weights import init_weights, synth_inputs

TINY = dict(io_shape=[4, 4, 2], x_d=1, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
            num_kernels_list=[8], cardinality_list=[2])
SMALL = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
             num_kernels_list=[16, 8], cardinality_list=[2, 2])


def test_planner_goldens():
    # SURVEY §8a A8 goldens (k=3)
    for S, want in [(4, [1]), (8, [1]), (14, [1, 2]), (16, [1, 2]), (28, [1, 2, 4]), (32, [1, 2, 4]),
                    (64, [1, 2, 4, 8]), (128, [1, 2, 4, 8, 16])]:
        p = plan_flow([S, S, 2], 1, [0], [1], [64], [2])
        assert [int(d) for d in p['dilations_list'][0]['channelwise']] == want
    p = plan_flow([28, 28, 2], 1, [0, 1, 0, 0], [3] * 4, [64, 64, 32, 32], [8, 8, 4, 4])
    assert p['scale_list'] == [1, 1, 2, 2]
    assert p['num_prev_factors_list'] == [0, 0, 1, 1]
    assert p['io_shape_list'] == [[28, 28, 2], [28, 28, 2], [14, 14, 4], [14, 14, 4]]
    assert [int(d) for d in p['dilations_list'][0]['checkerboard']] == [1, 2, 4]
    assert [int(d) for d in p['dilations_list'][2]['checkerboard']] == [1, 2]
    kinds = [L['type'] for L in p['layers']]
    assert kinds == ['coupling'] * 8 + ['squeeze', 'factor'] + ['coupling'] * 8
    cb, ch = p['layers'][0], p['layers'][2]
    assert (cb['h'], cb['w'], cb['c1'], cb['c2'], cb['nk'], cb['cat']) == (14, 14, 4, 4, 32, 56)
    assert (ch['h'], ch['w'], ch['c1'], ch['c2'], ch['nk'], ch['cat']) == (28, 28, 1, 1, 64, 112)
    assert [b['group_width'] for b in cb['branches']] == [4, 2, 1]     # Q4: cardinality not halved
    assert [b['group_width'] for b in ch['branches']] == [8, 4, 2]


def test_planner_quirk_q6():
    # defaults [64],[8] break on checkerboard layers once dilation 8 appears (F:396)
    with pytest.raises(AssertionError):
        plan_flow([64, 64, 6], 3, [0], [1], [64], [8])


@pytest.mark.parametrize("mask", [0, 2, 3])
def test_two_restatements_agree(mask):
    L = plan_coupling([8, 8, 3], mask, 2, 2, 16, 3, [1, 2])
    L['ln'] = True
    plan = {'layers': [L]}
    W = init_weights(plan, 'rand', seed=3)[0]
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, L['h'], L['w'], L['c1']))
    for net, is_A in (('A', True), ('b', False)):
        a = nets_np.st_net(x, W[net], L, is_A)
        Pt = {k: _t(v, torch.float64) for k, v in W[net].items()}
        b = st_net(_t(x, torch.float64), Pt, L, is_A).numpy()
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)


def test_invertibility_and_finiteness_at_init():
    o = FlowOracle(**SMALL)
    o.set_weights(init_weights(o.plan, 'init', seed=1))
    x = synth_inputs('noise:8x8x3', 3, seed=2)
    zy, ld, ps = o.call(x, 1)
    assert np.isfinite(zy).all() and np.isfinite(ps.numpy()).all()
    x2 = o.call(zy, -1)
    np.testing.assert_allclose(x2, x, atol=1e-5)


def test_invertibility_rand_weights_fp64():
    o = FlowOracle(**SMALL, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=1))
    x = synth_inputs('noise:8x8x3', 2, seed=5).astype(np.float64)
    zy, ld, ps = o.call(x, 1)
    np.testing.assert_allclose(o.call(zy, -1), x, atol=1e-9)
    # Q1: scalar log-det is the sum over layers of batch means == mean of per-sample sums
    np.testing.assert_allclose(float(ld), ps.numpy().mean(), rtol=1e-12)


def test_logdet_equals_autograd_jacobian():
    """log|det J| of the whole tiny flow from an autograd Jacobian == sum of A outputs (M:1323)."""
    o = FlowOracle(**TINY, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=7))
    x = synth_inputs('noise:4x4x2', 1, seed=9).astype(np.float64)
    _, _, ps = o.call(x, 1)

    # differentiable re-run through torch-only closed-form masks (gather/scatter by index)
    def f(flat):
        u = flat.reshape(1, 4, 4, 2)
        for li, L in enumerate(o.coupling):
            m, mc = L['mask'], L['mask_complement']
            idx = np.arange(32).reshape(1, 4, 4, 2)
            from oracle import masks_np as MM
            i1 = MM.mask(idx.astype(np.float64) + 1, m, True).astype(np.int64).ravel() - 1
            i2 = MM.mask(idx.astype(np.float64) + 1, mc, True).astype(np.int64).ravel() - 1
            shp1 = MM.mask(idx.astype(np.float64), m, True).shape
            u_flat = u.reshape(-1)
            u1c = u_flat[i1].reshape(shp1)
            A = st_net(u1c, o.W[li]['A'], L, True)
            b = st_net(u1c, o.W[li]['b'], L, False)
            v2 = torch.exp(A).reshape(-1) * u_flat[i2] + b.reshape(-1)
            out = u_flat.clone()
            out[i2] = v2
            u = out.reshape(1, 4, 4, 2)
        return u.reshape(-1)

    J = torch.autograd.functional.jacobian(f, torch.as_tensor(x.reshape(-1)))
    sign, logabs = np.linalg.slogdet(J.numpy())
    assert sign > 0
    np.testing.assert_allclose(logabs, float(ps[0]), rtol=1e-8)


def test_prior_and_loss_definition():
    o = FlowOracle(**TINY, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=7))
    x = synth_inputs('noise:4x4x2', 4, seed=11).astype(np.float64)
    four, ps = o.log_loss(x)
    zy = ps['zy']
    ll_z = (-0.5 * zy[..., :1] ** 2 - 0.5 * math.log(2 * math.pi)).sum(axis=(1, 2, 3))
    ll_y = -100.0 * np.abs(zy[..., 1:] - x[..., 1:]).sum(axis=(1, 2, 3))
    np.testing.assert_allclose(ps['ll_z'], ll_z, rtol=1e-12)
    np.testing.assert_allclose(ps['ll_y'], ll_y, rtol=1e-12)
    np.testing.assert_allclose(four[0], -((ll_z + ll_y).mean() + ps['logdet'].mean()), rtol=1e-12)
    np.testing.assert_allclose(four[0], four[1] + four[2] + four[3], rtol=1e-12)


def test_init_state_A_small():
    o = FlowOracle(**SMALL)
    o.set_weights(init_weights(o.plan, 'init', seed=1))
    x = synth_inputs('noise:8x8x3', 2, seed=2)
    L = o.coupling[0]
    from oracle import masks_np as MM
    A, b = o._nets(0, MM.mask(x, L['mask'], True))
    assert float(A.abs().max()) < 1.0
