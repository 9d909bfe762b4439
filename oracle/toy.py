"""Oracle: NumPy restatement of the toy dense cINN (BASELINE config 1).

Follows /root/reference/TOYcINN_make_model.py:
  * coupling_layer (dense s/t MLPs)  T:29-97   (Dense + LeakyReLU(0.3); A ends in tanh, no learned scale)
  * masks                            T:154-166 (u1 index sets {0},{1},{2},{0,1},{0,2},{1,2}; u2 = complement)
  * cINN_affine.call                 T:248-402 (direction -1: x->z, layers n-1..0, v2 = exp(A)u2+b,
                                               per-sample log_detJ += log(prod exp A); direction +1 inverse)
  * log_loss                         T:404-451
Weights per coupling layer: {'A': [(W0,b0),...], 'b': [...]}, Keras Dense layout (in, out).
TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import math

import numpy as np

MASK_1 = {0: [0], 1: [1], 2: [2], 3: [0, 1], 4: [0, 2], 5: [1, 2]}     # T:154-159
MASK_2 = {0: [1, 2], 1: [0, 2], 2: [0, 1], 3: [2], 4: [1], 5: [0]}     # T:160-165


def toy_init_weights(num_coupling_layers, intermediate_dims, num_layers, seed=0, scale=None):
    """Glorot-uniform Dense kernels (the Keras default actually used, T:52-93; `init` is unused, T:138),
    zero biases.  `scale` multiplies the limit to get a livelier 'trained-like' set."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(num_coupling_layers):
        d1, d2 = len(MASK_1[i % 6]), len(MASK_2[i % 6])
        dims = [d1] + [intermediate_dims] * (num_layers + 1) + [d2]
        entry = {}
        for net in ('A', 'b'):
            layers = []
            for a, b in zip(dims[:-1], dims[1:]):
                lim = math.sqrt(6.0 / (a + b)) * (scale or 1.0)
                W = rng.uniform(-lim, lim, (a, b)).astype(np.float32)
                bias = (np.zeros(b, np.float32) if scale is None
                        else (0.1 * rng.standard_normal(b)).astype(np.float32))
                layers.append((W, bias))
            entry[net] = layers
        out.append(entry)
    return out


def mlp(x, layers, final_tanh):
    h = x
    for i, (W, b) in enumerate(layers):
        h = h @ W + b
        if i < len(layers) - 1:
            h = np.where(h > 0, h, 0.3 * h)
    return np.tanh(h) if final_tanh else h


class ToyOracle:
    def __init__(self, io_shape, x_d, num_coupling_layers, weights, mask_indices=None, dtype=np.float32):
        assert io_shape == 3
        self.n, self.x_d, self.dtype = num_coupling_layers, x_d, dtype
        self.mask_indices = (list(mask_indices) if mask_indices is not None
                             else list(range(num_coupling_layers)))
        self.W = [{net: [(np.asarray(W, dtype), np.asarray(b, dtype)) for W, b in w[net]]
                   for net in ('A', 'b')} for w in weights]
        self.lambda_y = 100

    def call(self, u, direction=-1):
        u = np.asarray(u, self.dtype).copy()
        log_detJ = np.zeros(u.shape[0], self.dtype)
        for i in list(range(self.n))[::direction]:          # T:295
            j = self.mask_indices[i]
            m1, m2 = MASK_1[j % 6], MASK_2[j % 6]
            u1, u2 = u[:, m1], u[:, m2]                      # T:311-319 (0/1 matvec == gather)
            A = mlp(u1, self.W[j]['A'], True)
            b = mlp(u1, self.W[j]['b'], False)
            eA = np.exp(A)
            if direction == 1:
                t = (1.0 / eA) * (u2 - b)                    # T:369-375
            else:
                t = eA * u2 + b                              # T:379-381
                log_detJ = log_detJ + np.log(np.prod(eA, axis=1))   # T:386-387 (Q12)
            out = np.zeros_like(u)
            out[:, m1] = u1                                  # T:366, T:395 (transpose matvec == scatter)
            out[:, m2] = t
            u = out
        return u, (log_detJ if direction == -1 else 0)

    def log_loss(self, xy):
        xy = np.asarray(xy, self.dtype)
        x_d = self.x_d
        y_prime = xy[:, x_d:]
        zy, ld = self.call(xy, -1)
        z, y = zy[:, :x_d], zy[:, x_d:]
        ll_z = -0.5 * (z * z).sum(1) - 0.5 * x_d * math.log(2 * math.pi)
        ll_y = -self.lambda_y * np.abs(y - y_prime).sum(1)
        ll = (ll_z + ll_y + ld).mean()
        return (float(-ll), float(-ll_z.mean()), float(-ll_y.mean()), float(-ld.mean())), \
            {'ll_z': ll_z, 'll_y': ll_y, 'logdet': ld, 'zy': zy}
