"""Oracle: gradients of cFlow.log_loss w.r.t. every trainable variable, by torch autograd.

The reference trains with `tf.GradientTape` over `log_loss` (conv_cINN_make_model.py M:1863-1871); this is
the same computation restated end-to-end in differentiable torch ops (the s/t nets are the functions of
`flow_torch`; masks use the closed-form gather/scatter that `tests/test_oracle_masks.py` pins against the
literal transcription).  Also restates one Keras Adam step (lr 3e-4, beta 0.9/0.999, eps 1e-7; C:567, P:130).
TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import math

import numpy as np
import torch

from .flow_torch import st_net


def _compressed_index(H, W, D, m):
    """flat indices (into an H*W*D sample) of mask(u, m, compress=True), in compressed order"""
    idx = torch.arange(H * W * D).reshape(H, W, D)
    if m in (0, 1):
        a = m
        return torch.cat([idx[0::2, a::2, :], idx[1::2, (1 - a)::2, :]], dim=-1).reshape(-1)
    return idx[..., (m - 2)::2].reshape(-1)


def loss_and_grads(oracle, xy, dtype=torch.float64):
    """oracle: FlowOracle with weights set.  Returns (loss4 floats, grads) where grads mirrors the weights
    structure: list over coupling layers of {'A': {name: ndarray}, 'b': {...}}."""
    plan = oracle.plan
    H0, W0, D0 = plan['io_shape']
    x_d, lam = oracle.x_d, oracle.lambda_y
    Wt = [{net: {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in w[net].items()}
           for net in ('A', 'b')} for w in oracle.W]
    xy_t = torch.as_tensor(np.asarray(xy)).to(dtype)
    B = xy_t.shape[0]
    buf = xy_t.reshape(B, H0, W0, D0)
    logdet = torch.zeros(B, dtype=dtype)
    level, ci = 0, 0
    for L in oracle.layers:
        if L['type'] == 'squeeze':
            level += 1
            continue
        if L['type'] == 'factor':
            continue
        S = 2 ** level
        H, Wd, D = H0 // S, W0 // S, D0 * S
        # strided active view (tests/test_oracle_masks.py::test_active_tensor_is_a_strided_view)
        act = buf[:, S - 1::S, :, :].reshape(B, H, Wd, D)
        flat = act.reshape(B, -1)
        i1 = _compressed_index(H, Wd, D, L['mask'])
        i2 = _compressed_index(H, Wd, D, L['mask_complement'])
        u1c = flat[:, i1].reshape(B, L['h'], L['w'], L['c1'])
        A = st_net(u1c, Wt[ci]['A'], L, True).reshape(B, -1)
        t = st_net(u1c, Wt[ci]['b'], L, False).reshape(B, -1)
        v2 = torch.exp(A) * flat[:, i2] + t
        new_flat = flat.clone()
        new_flat[:, i2] = v2
        logdet = logdet + A.sum(dim=1)
        new_act = new_flat.reshape(B, H, Wd, D).reshape(B, H, W0, D0)
        nb = buf.clone()
        nb[:, S - 1::S, :, :] = new_act
        buf = nb
        ci += 1
    zy = buf
    z, y = zy[..., :x_d], zy[..., x_d:]
    ll_z = (-0.5 * (z * z).sum(-1) - 0.5 * x_d * math.log(2 * math.pi)).sum(dim=(1, 2))
    ll_y = -lam * (y - xy_t[..., x_d:]).abs().sum(dim=(1, 2, 3))
    loss = -((ll_z + ll_y).mean() + logdet.mean())
    leaves = [p for w in Wt for net in ('A', 'b') for p in w[net].values()]
    gs = torch.autograd.grad(loss, leaves, allow_unused=True)
    it = iter(gs)
    grads = []
    for w in Wt:
        e = {}
        for net in ('A', 'b'):
            e[net] = {}
            for k, p in w[net].items():
                g = next(it)
                e[net][k] = (torch.zeros_like(p) if g is None else g).numpy()
        grads.append(e)
    four = (float(loss.detach()), float(-ll_z.detach().mean()), float(-ll_y.detach().mean()), float(-logdet.detach().mean()))
    return four, grads


def adam_step(p, g, m, v, step, lr=3e-4, b1=0.9, b2=0.999, eps=1e-7):
    """One Keras Adam update (non-amsgrad): returns (p, m, v).  step counts from 1."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    alpha = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    p = p - alpha * m / (np.sqrt(v) + eps)
    return p, m, v
