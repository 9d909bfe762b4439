"""Entry points of the reference surface that round 1 left as stubs: cINN_affine.train_step (T:453-482),
dilated_residual_block on a tensor (F:501-627), coupling_layer.coupling_function (M:1076-1213)."""
import numpy as np
import pytest
import torch

from oracle.planner import plan_coupling
from oracle.weights import init_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _toy(dev, n, width, num_layers, seed=3):
    from arl_conditional_normalizing_flows_b200.TOYcINN_make_model import cINN_affine
    from oracle.toy import toy_init_weights
    W = toy_init_weights(n, width, num_layers, seed=seed, scale=1.0)
    order = list(np.random.default_rng(1).permutation(n))
    m = cINN_affine(3, 2, n, width, num_layers, None, mask_indices=order, device=dev)
    for j, cl in enumerate(m.coupling_layers_list):
        flat = []
        for net in ('b', 'A'):
            for Wm, bv in W[j][net]:
                flat += [Wm, bv]
        cl.set_weights(flat)
    return m, W, order


@pytest.mark.parametrize("width,num_layers,n,B", [(32, 6, 24, 1000), (16, 2, 12, 77), (8, 0, 6, 5), (64, 1, 6, 130)])
def test_toy_gradients_match_autograd_oracle(dev, width, num_layers, n, B):
    from oracle.toy_grad_torch import toy_loss_and_grads
    m, W, order = _toy(dev, n, width, num_layers)
    rng = np.random.default_rng(0)
    xy = np.concatenate([rng.standard_normal((B, 2)), np.where(rng.uniform(size=(B, 1)) < 0.5, -1.0, 1.0)], 1)
    four_want, grads_want = toy_loss_and_grads(W, xy, 2, order)
    four, _ = m.loss_and_grad(torch.from_numpy(xy.astype(np.float32)).to(dev))
    np.testing.assert_allclose([float(t) for t in four], four_want, rtol=1e-4, atol=1e-5)
    got = m.grad_views()
    worst = 0.0
    for j in range(n):
        for net in ('A', 'b'):
            for (gw, gb), (ww, wb) in zip(got[j][net], grads_want[j][net]):
                for a, b_ in ((gw, ww), (gb, wb)):
                    # tolerance: 2e-3 of the tensor's largest entry (fp32 kernels, layer inputs recovered by the inverse
                    # law instead of stored; the conv model's gradient gate, tests/test_gpu_train.py)
                    err = float(np.abs(a - b_).max() / max(np.abs(b_).max(), 1e-12))
                    worst = max(worst, err)
                    assert err < 2e-3, (j, net, err)
    assert worst > 0.0


def test_toy_train_step_updates_like_adam(dev):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import Adam
    from oracle.grad_torch import adam_step
    from oracle.toy_grad_torch import toy_loss_and_grads
    m, W, order = _toy(dev, 12, 16, 2)
    rng = np.random.default_rng(5)
    xy = np.concatenate([rng.standard_normal((200, 2)), np.where(rng.uniform(size=(200, 1)) < 0.5, -1.0, 1.0)], 1)
    four_want, grads_want = toy_loss_and_grads(W, xy, 2, order)
    m.compile(optimizer=Adam(3e-4))
    logs = m.train_step(torch.from_numpy(xy.astype(np.float32)).to(dev))
    assert set(logs) == {'loss', 'z_loss', 'y_loss', 'detJ_loss'}          # T:478-482
    np.testing.assert_allclose(logs['loss'], four_want[0], rtol=1e-4)
    for j, cl in enumerate(m.coupling_layers_list):                        # the layers' get_weights() see the update
        new = cl.get_weights()
        k = 0
        for net in ('b', 'A'):
            for (w0, b0), (gw, gb) in zip(W[j][net], grads_want[j][net]):
                for p0, g in ((w0, gw), (b0, gb)):
                    want, _, _ = adam_step(np.asarray(p0, np.float64), np.asarray(g, np.float64), 0.0, 0.0, 1)
                    # one Adam step moves every weight by ~lr * sign(g): compare the moved weights where |g| is not tiny
                    mask = np.abs(g) > 1e-3 * np.abs(g).max()
                    np.testing.assert_allclose(np.asarray(new[k], np.float64)[mask], want[mask], rtol=0, atol=2e-5)
                    k += 1
    # a few more steps decrease the loss on the same batch
    first = logs['loss']
    for _ in range(20):
        m.loss_tracker.reset_state()
        logs = m.train_step(torch.from_numpy(xy.astype(np.float32)).to(dev))
    assert logs['loss'] < first


@pytest.mark.parametrize("h,w,nk,card,dil,ln", [(14, 14, 32, 4, [1, 2], True), (28, 28, 64, 8, [1, 2, 4], True),
                                               (7, 5, 16, 2, [1, 2], True), (12, 12, 16, 4, [1], False)])
def test_dilated_residual_block_on_a_tensor(dev, h, w, nk, card, dil, ln):
    """F:501-627 called the way the reference calls it (M:1123-1130), against the oracle's restatement."""
    from arl_conditional_normalizing_flows_b200.conv_cINN_base_functions import dilated_residual_block
    from oracle.flow_torch import dilated_residual_block as oracle_block, _t
    L = plan_coupling([2 * h, 2 * w, 1], 0, 1, card, 2 * nk, 3, dil)
    L['ln'] = ln
    Wfull = init_weights({'layers': [L]}, 'rand', seed=21, ln=ln)[0]['b']
    Wblk = {k[4:]: v for k, v in Wfull.items() if k.startswith('rb0.')}
    rng = np.random.default_rng(2)
    y = rng.standard_normal((3, h, w, nk)).astype(np.float32)
    got = dilated_residual_block(torch.from_numpy(y).to(dev), nk, nk, _which_dilations=dil, ksize=(3, 3), cardinality=card,
                                 ln=ln, weights=Wblk)
    P = {k: _t(v, torch.float64) for k, v in Wfull.items()}
    with torch.no_grad():
        want = oracle_block(_t(y.astype(np.float64), torch.float64), P, 0, L).numpy()
    err = np.abs(got.cpu().numpy() - want).max() / np.abs(want).max()
    assert err < 1e-4, err
    # weights=None builds a fresh block like the Keras builder; weights={} hands the fresh parameters back
    fresh = {}
    out = dilated_residual_block(torch.from_numpy(y).to(dev), nk, nk, _which_dilations=dil, ksize=(3, 3), cardinality=card,
                                 ln=ln, weights=fresh)
    assert out.shape == got.shape and torch.isfinite(out).all() and 'pw1.kernel' in fresh
    with pytest.raises(NotImplementedError):
        dilated_residual_block(torch.from_numpy(y).to(dev), nk, nk, _strides=(2, 2))
    with pytest.raises(NotImplementedError):
        dilated_residual_block(torch.from_numpy(y).to(dev), nk // 2, nk)


def test_coupling_function_returns_the_two_networks(dev):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
    from oracle import masks_np
    from oracle.flow_torch import st_net, _t
    shape, m, R, card, nk, dil = [14, 14, 4], 2, 2, 4, 32, [1, 2]
    L = plan_coupling(shape, m, R, card, nk, 3, dil)
    L['ln'] = True
    W = init_weights({'layers': [L]}, 'rand', seed=8)[0]
    layer = coupling_layer(shape, m, R, card, nk, 3, None, LAYER_NORM=True, which_dilations=dil, device=dev)
    layer.set_weights(W)
    model_A, model_b = layer.coupling_function()
    u = np.random.default_rng(1).standard_normal((2, *shape)).astype(np.float32)
    u1c = masks_np.mask(u.astype(np.float64), m, True)
    with torch.no_grad():
        A = st_net(_t(u1c, torch.float64), {k: _t(v, torch.float64) for k, v in W['A'].items()}, L, True).numpy()
        b = st_net(_t(u1c, torch.float64), {k: _t(v, torch.float64) for k, v in W['b'].items()}, L, False).numpy()
    u1c_t = torch.from_numpy(np.ascontiguousarray(u1c, dtype=np.float32)).to(dev)
    assert np.abs(model_A(u1c_t).cpu().numpy() - A).max() / np.abs(A).max() < 1e-4
    assert np.abs(model_b(u1c_t).cpu().numpy() - b).max() / np.abs(b).max() < 1e-4
    assert set(model_A.get_weights()) == set(W['A']) and set(model_b.get_weights()) == set(W['b'])


@pytest.mark.parametrize("B", [1, 7, 300])
def test_call_forward_scalar_is_the_batch_mean_of_the_per_sample_logdets(dev, B):
    """cFlow.call(xy, +1) returns (zy, scalar): the scalar (M:1325-1326, Q1) comes out of the same launch as the per-sample
    vector (logdet [B + 1] through the C-ABI), not out of an eager reduction."""
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    m = cFlow(io_shape=[8, 8, 2], x_d=1, squeeze_factor_block_list=[0, 1], ResNeXt_block_list=[1, 1],
              num_kernels_list=[16, 16], cardinality_list=[2, 2], device=dev)
    m.randomize_weights(seed=1)
    xy = torch.randn(B, 8, 8, 2, generator=torch.Generator().manual_seed(B)).to(dev)
    zy, ld = m(xy, 1)
    ps = m.last_logdet_per_sample
    assert ps.shape == (B,) and ld.dim() == 0
    want = ps.double().mean().item()
    assert abs(float(ld) - want) <= 1e-6 * max(1.0, abs(want))


def test_log_loss_and_sample_on_two_streams_equals_the_two_calls():
    """cFlow.log_loss_and_sample issues the two independent halves of an evaluation + sampling step on two CUDA streams
    (separate workspaces): same kernels, bit-identical results, repeated calls included."""
    import torch
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    dev = torch.device("cuda:0")
    cfg = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[1, 1, 1, 1],
               num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
    m = cFlow(**cfg, device=dev)
    m.randomize_weights(seed=3)
    g = torch.Generator().manual_seed(0)
    for B in (33, 8):
        x = torch.randn(B, 28, 28, 2, generator=g).to(dev)
        z = torch.randn(B, 28, 28, 2, generator=g).to(dev)
        want_four = [float(t) for t in m.log_loss(x)]
        want_ld = m.last_per_sample['logdet'].clone()
        want_s = m(z, -1).clone()
        for _ in range(3):
            four, s = m.log_loss_and_sample(x, z)
            torch.cuda.synchronize()
            assert [float(t) for t in four] == want_four
            assert torch.equal(m.last_per_sample['logdet'], want_ld)
            assert torch.equal(s, want_s)
