"""Oracle: gradient of the toy log_loss by torch autograd (fp64) over a torch restatement of oracle/toy.py.

Follows /root/reference/TOYcINN_make_model.py: cINN_affine.call T:248-402 (direction -1), log_loss T:404-451,
train_step T:453-482 (tf.GradientTape over log_loss w.r.t. every Dense kernel and bias).
TEST INFRASTRUCTURE (see oracle/__init__.py): the restatement is checked against oracle/toy.py in
tests/test_oracle_toy_grad.py before it is trusted as the gradient reference.
"""
import math

import numpy as np
import torch

from .toy import MASK_1, MASK_2


def _mlp(x, layers, final_tanh):
    h = x
    for i, (W, b) in enumerate(layers):
        h = h @ W + b
        if i < len(layers) - 1:
            h = torch.nn.functional.leaky_relu(h, 0.3)
    return torch.tanh(h) if final_tanh else h


def toy_loss_and_grads(weights, xy, x_d, mask_indices=None, lambda_y=100.0):
    """weights: [{'A': [(W, b), ...], 'b': [...]}, ...] (numpy); returns ((loss, z_loss, y_loss, detJ_loss),
    grads with the same nesting as `weights`)."""
    n = len(weights)
    order = list(mask_indices) if mask_indices is not None else list(range(n))
    Wt = [{net: [(torch.tensor(np.asarray(W, np.float64), requires_grad=True),
                  torch.tensor(np.asarray(b, np.float64), requires_grad=True)) for W, b in w[net]]
           for net in ('A', 'b')} for w in weights]
    u = torch.tensor(np.asarray(xy, np.float64))
    y_prime = u[:, x_d:]
    ld = torch.zeros(u.shape[0], dtype=torch.float64)
    for i in list(range(n))[::-1]:                         # T:295, direction -1
        j = order[i]
        m1, m2 = MASK_1[j % 6], MASK_2[j % 6]
        u1, u2 = u[:, m1], u[:, m2]
        A = _mlp(u1, Wt[j]['A'], True)
        b = _mlp(u1, Wt[j]['b'], False)
        eA = torch.exp(A)
        t = eA * u2 + b                                    # T:379-381
        ld = ld + torch.log(torch.prod(eA, dim=1))         # T:386-387
        out = torch.zeros_like(u)
        out[:, m1] = u1
        out[:, m2] = t
        u = out
    z, y = u[:, :x_d], u[:, x_d:]
    ll_z = -0.5 * (z * z).sum(1) - 0.5 * x_d * math.log(2 * math.pi)
    ll_y = -lambda_y * (y - y_prime).abs().sum(1)
    loss = -(ll_z + ll_y + ld).mean()
    loss.backward()
    grads = [{net: [(W.grad.numpy().copy(), b.grad.numpy().copy()) for W, b in w[net]] for net in ('A', 'b')} for w in Wt]
    return (float(loss.detach()), float(-ll_z.mean().detach()), float(-ll_y.mean().detach()), float(-ld.mean().detach())), grads
