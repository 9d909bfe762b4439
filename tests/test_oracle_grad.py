"""The gradient oracle (oracle/grad_torch.py: torch autograd over the restated log_loss, M:1863-1871) is
pinned here against (i) the forward oracle it must reproduce, (ii) central finite differences of that
forward oracle, (iii) a hand-computed Adam step (keras semantics, eps outside the sqrt)."""
import numpy as np
import torch

from oracle.flow_torch import FlowOracle
from oracle.grad_torch import adam_step, loss_and_grads
from oracle.weights import init_weights

SMALL = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
             num_kernels_list=[16, 8], cardinality_list=[2, 2])


def _mk(seed=1):
    o = FlowOracle(**SMALL, dtype=torch.float64)
    W = init_weights(o.plan, 'rand', seed=seed)
    o.set_weights(W)
    return o, W


def test_autograd_restatement_reproduces_log_loss():
    o, _ = _mk()
    xy = np.random.default_rng(0).standard_normal((3, 8, 8, 3))
    want, _ = o.log_loss(xy)
    four, grads = loss_and_grads(o, xy)
    np.testing.assert_allclose(four, want, rtol=1e-10)
    assert len(grads) == len(o.W)
    for g, w in zip(grads, o.W):
        for net in ('A', 'b'):
            assert set(g[net]) == set(w[net])


def test_gradients_match_finite_differences():
    o, W = _mk()
    xy = np.random.default_rng(0).standard_normal((2, 8, 8, 3))
    _, grads = loss_and_grads(o, xy)
    rng = np.random.default_rng(5)
    probes = [(0, 'A', 'stem.kernel'), (0, 'A', 'tanh_scale'), (1, 'b', 'rb0.ln2.gamma'), (3, 'A', 'rb1.pw2.kernel'),
              (5, 'b', 'head.bias'), (6, 'A', 'rb0.ln1.beta'), (2, 'A', 'rb0.gc.d1.g1.kernel'), (7, 'b', 'lnf.gamma')]
    for li, net, name in probes:
        assert name in W[li][net], (name, sorted(W[li][net])[:8])
        base = np.array(W[li][net][name], np.float64)
        idx = tuple(rng.integers(0, s) for s in base.shape)
        h = 1e-5
        vals = []
        for sgn in (+1, -1):
            Wp = [{n: dict(d) for n, d in w.items()} for w in W]
            pert = base.copy()
            pert[idx] += sgn * h
            Wp[li][net][name] = pert
            o.set_weights(Wp)
            vals.append(o.log_loss(xy)[0][0])
        o.set_weights(W)
        fd = (vals[0] - vals[1]) / (2 * h)
        g = np.asarray(grads[li][net][name])[idx]
        assert abs(fd - g) <= 1e-5 * max(1.0, abs(g)) + 1e-6, (li, net, name, fd, g)


def test_adam_step_known_answer():
    p, g = np.array([1.0, -2.0]), np.array([0.5, -0.25])
    p1, m1, v1 = adam_step(p, g, np.zeros(2), np.zeros(2), 1, lr=3e-4)
    # first step: m_hat = g, v_hat = g^2 -> update = lr * g / (|g| + eps*sqrt(1-b2)) ~ lr * sign(g)
    np.testing.assert_allclose(m1, 0.1 * g)
    np.testing.assert_allclose(v1, 0.001 * g * g)
    want = p - 3e-4 * g / (np.abs(g) + 1e-7 / np.sqrt(1 - 0.999))
    np.testing.assert_allclose(p1, want, rtol=0, atol=1e-12)
