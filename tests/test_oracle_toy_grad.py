"""The torch restatement behind the toy gradient oracle reproduces the NumPy toy oracle's loss, and its autograd
gradient matches central finite differences of that NumPy oracle (so the gradient reference is pinned to the same
restatement the forward tests use)."""
import numpy as np

from oracle.toy import ToyOracle, toy_init_weights
from oracle.toy_grad_torch import toy_loss_and_grads


def test_torch_restatement_and_gradient():
    n, width, L = 6, 8, 1
    W = toy_init_weights(n, width, L, seed=2, scale=1.0)
    rng = np.random.default_rng(0)
    xy = np.concatenate([rng.standard_normal((7, 2)), np.where(rng.uniform(size=(7, 1)) < 0.5, -1.0, 1.0)], 1)
    order = list(rng.permutation(n))
    four, grads = toy_loss_and_grads(W, xy, 2, order)
    want, _ = ToyOracle(3, 2, n, W, mask_indices=order, dtype=np.float64).log_loss(xy)
    np.testing.assert_allclose(four, want, rtol=1e-12)
    eps = 1e-6
    for (j, net, li, idx) in [(0, 'A', 0, (0, 3)), (3, 'b', 1, (2, 5)), (5, 'A', 2, (4, 0)), (2, 'b', 0, (0, 1))]:
        Wp = [{k: [(w.astype(np.float64).copy(), b.astype(np.float64).copy()) for w, b in v[k]] for k in ('A', 'b')} for v in W]
        Wm = [{k: [(w.astype(np.float64).copy(), b.astype(np.float64).copy()) for w, b in v[k]] for k in ('A', 'b')} for v in W]
        Wp[j][net][li][0][idx] += eps
        Wm[j][net][li][0][idx] -= eps
        fp = ToyOracle(3, 2, n, Wp, mask_indices=order, dtype=np.float64).log_loss(xy)[0][0]
        fm = ToyOracle(3, 2, n, Wm, mask_indices=order, dtype=np.float64).log_loss(xy)[0][0]
        np.testing.assert_allclose(grads[j][net][li][0][idx], (fp - fm) / (2 * eps), rtol=2e-5, atol=1e-7)
