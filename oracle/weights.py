"""Oracle: seeded weight sets for parity tests and benchmarks.

`init_weights(plan, kind, seed)` returns, for every coupling layer in `layers_list` order,
{'A': {...}, 'b': {...}} with Keras-shaped arrays (HWIO kernels, flat LN vectors):

  kind='init'  reference initial state: Orthogonal(gain 0.1) kernels
               (conv_cINN_make_model.py M:1442, Keras semantics: QR of a normal matrix flattened
               to (kh*kw*cin, cout)), zero biases, gamma=1, beta=0, tanh scale w=1 (M:109-112).
  kind='rand'  "trained-like": kernels N(0, 1/fan_in), biases N(0, 0.1), gamma U[0.5,1.5],
               beta N(0, 0.1), w U[0.5,1.5]; exercises exp / log-det away from zero.
TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np


def orthogonal(rng, shape, gain):
    rows = int(np.prod(shape[:-1]))
    cols = int(shape[-1])
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (gain * q.reshape(shape)).astype(np.float32)


def param_specs(L):
    """Ordered (name, shape, role) for one s/t net of coupling-layer plan L (A-net adds tanh_scale)."""
    k, nk, c1, c2, h, w = L['ksize'], L['nk'], L['c1'], L['c2'], L['h'], L['w']
    ln = L.get('ln', True)
    s = [('stem.kernel', (k, k, c1, nk), 'kernel'), ('stem.bias', (nk,), 'bias')]
    for r in range(L['R']):
        if ln:
            s += [(f'rb{r}.ln1.gamma', (h * w * nk,), 'gamma'), (f'rb{r}.ln1.beta', (h * w * nk,), 'beta')]
        s += [(f'rb{r}.pw1.kernel', (1, 1, nk, nk), 'kernel'), (f'rb{r}.pw1.bias', (nk,), 'bias')]
        if ln:
            s += [(f'rb{r}.ln2.gamma', (h * w * nk,), 'gamma'), (f'rb{r}.ln2.beta', (h * w * nk,), 'beta')]
        for br in L['branches']:
            d = br['dilation']
            if L['cardinality'] == 1:
                s += [(f'rb{r}.gc.d{d}.g0.kernel', (k, k, nk, br['channels']), 'kernel'),
                      (f'rb{r}.gc.d{d}.g0.bias', (br['channels'],), 'bias')]
            else:
                g = br['group_width']
                for j in range(L['cardinality']):
                    s += [(f'rb{r}.gc.d{d}.g{j}.kernel', (k, k, g, g), 'kernel'),
                          (f'rb{r}.gc.d{d}.g{j}.bias', (g,), 'bias')]
        if ln:
            s += [(f'rb{r}.ln3.gamma', (h * w * L['cat'],), 'gamma'), (f'rb{r}.ln3.beta', (h * w * L['cat'],), 'beta')]
        s += [(f'rb{r}.pw2.kernel', (1, 1, L['cat'], nk), 'kernel'), (f'rb{r}.pw2.bias', (nk,), 'bias')]
    if ln:
        s += [('lnf.gamma', (h * w * nk,), 'gamma'), ('lnf.beta', (h * w * nk,), 'beta')]
    s += [('head.kernel', (k, k, nk, c2), 'kernel'), ('head.bias', (c2,), 'bias')]
    return s


def init_weights(plan, kind='rand', seed=0, ln=True):
    rng = np.random.default_rng(seed)
    out = []
    for L in plan['layers']:
        if L['type'] != 'coupling':
            continue
        L = dict(L, ln=ln)
        entry = {}
        for net in ('A', 'b'):
            P = {}
            for name, shape, role in param_specs(L):
                if kind == 'init':
                    if role == 'kernel':
                        P[name] = orthogonal(rng, shape, 0.1)
                    elif role == 'gamma':
                        P[name] = np.ones(shape, np.float32)
                    else:
                        P[name] = np.zeros(shape, np.float32)
                else:
                    if role == 'kernel':
                        fan_in = int(np.prod(shape[:-1]))
                        P[name] = (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
                    elif role == 'gamma':
                        P[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
                    else:
                        P[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
            if net == 'A':
                P['tanh_scale'] = (np.float32(1.0) if kind == 'init'
                                   else np.float32(rng.uniform(0.5, 1.5)))
            entry[net] = P
        out.append(entry)
    return out


def synth_inputs(cfg, B, seed=0):
    """Synthetic xy batches of SURVEY §8(d).  cfg in {'cfg2','cfg3','cfg4','cfg5','noise:<H>x<W>x<D>'}."""
    rng = np.random.default_rng(seed)
    if cfg == 'cfg4':
        # super-resolution: hr ~ U[0,1) 64x64x3, y = 8x8 block means repeated to 64x64 (up^3(down^3(hr)), F:74-164),
        # x = hr - y (RESIDUAL, F:252-253), xy = concat(x, y) with 2 % noise (C:312-315)
        from .data_np import preprocess_SR
        hr = rng.uniform(0, 1, (B, 64, 64, 3)).astype(np.float32)
        xy = preprocess_SR(hr, levels=(0, 3)).astype(np.float64)
        return (0.98 * xy + 0.02 * rng.standard_normal(xy.shape)).astype(np.float32)
    if cfg == 'cfg5':
        return rng.standard_normal((B, 128, 128, 4)).astype(np.float32)   # pre-training on noise (P:102-115)
    if cfg in ('cfg2', 'cfg3'):
        H, W, xd = (28, 28, 1) if cfg == 'cfg2' else (32, 32, 3)
        img = 0.98 * rng.uniform(0, 1, (B, H, W, xd)) + 0.02 * rng.standard_normal((B, H, W, xd))
        label = rng.integers(0, 10) / 9.0
        lab = 0.98 * np.full((B, H, W, 1), label) + 0.02 * rng.standard_normal((B, H, W, 1))
        return np.concatenate([img, lab], axis=-1).astype(np.float32)
    if cfg.startswith('noise:'):
        H, W, D = (int(t) for t in cfg.split(':')[1].split('x'))
        return rng.standard_normal((B, H, W, D)).astype(np.float32)
    raise ValueError(cfg)
