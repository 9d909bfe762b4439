"""Oracle: second, independent NumPy restatement of the s/t-network arithmetic.

Written without torch (explicit zero padding + shifted slices) so that the two
restatements (`flow_torch` and this file) pin each other on the Keras semantics the
reference relies on: Conv2D 'same' padding with dilation (F:402-408, M:1114, M:1150),
LeakyReLU 0.3 (F:345), flattened-sample LayerNorm eps 1e-3 (F:350-360), grouped
channel slicing (F:397-411), residual add (F:625), w*tanh (M:1198, M:114-116).
Runs in float64 by default; small shapes only.  TEST INFRASTRUCTURE.
"""
import numpy as np


def conv2d_same(x, kernel, bias, dilation=1):
    B, H, W, Cin = x.shape
    k = kernel.shape[0]
    pad = dilation * (k - 1) // 2
    xp = np.zeros((B, H + 2 * pad, W + 2 * pad, Cin), x.dtype)
    xp[:, pad:pad + H, pad:pad + W, :] = x
    out = np.zeros((B, H, W, kernel.shape[3]), x.dtype)
    for ky in range(k):
        for kx in range(k):
            patch = xp[:, ky * dilation:ky * dilation + H, kx * dilation:kx * dilation + W, :]
            out += patch @ kernel[ky, kx]
    return out + bias


def common_layers(y, gamma, beta, ln=True):
    y = np.where(y > 0, y, 0.3 * y)
    if ln:
        B = y.shape[0]
        f = y.reshape(B, -1)
        mean = f.mean(axis=1, keepdims=True)
        var = ((f - mean) ** 2).mean(axis=1, keepdims=True)
        f = (f - mean) / np.sqrt(var + 1e-3) * gamma + beta
        y = f.reshape(y.shape)
    return y


def st_net(u1c, P, L, is_A, dtype=np.float64):
    P = {k: np.asarray(v, dtype) for k, v in P.items()}
    ln = L.get('ln', True)
    x = conv2d_same(np.asarray(u1c, dtype), P['stem.kernel'], P['stem.bias'])
    for r in range(L['R']):
        sc = x
        y = common_layers(x, P.get(f'rb{r}.ln1.gamma'), P.get(f'rb{r}.ln1.beta'), ln)
        y = conv2d_same(y, P[f'rb{r}.pw1.kernel'], P[f'rb{r}.pw1.bias'])
        y = common_layers(y, P.get(f'rb{r}.ln2.gamma'), P.get(f'rb{r}.ln2.beta'), ln)
        outs = []
        for br in L['branches']:
            d = br['dilation']
            if L['cardinality'] == 1:
                outs.append(conv2d_same(y, P[f'rb{r}.gc.d{d}.g0.kernel'], P[f'rb{r}.gc.d{d}.g0.bias'], d))
                continue
            g = br['group_width']
            for j in range(L['cardinality']):
                outs.append(conv2d_same(y[..., j * g:(j + 1) * g], P[f'rb{r}.gc.d{d}.g{j}.kernel'],
                                        P[f'rb{r}.gc.d{d}.g{j}.bias'], d))
        y = np.concatenate(outs, axis=-1)
        y = common_layers(y, P.get(f'rb{r}.ln3.gamma'), P.get(f'rb{r}.ln3.beta'), ln)
        y = conv2d_same(y, P[f'rb{r}.pw2.kernel'], P[f'rb{r}.pw2.bias'])
        x = sc + y
    x = common_layers(x, P.get('lnf.gamma'), P.get('lnf.beta'), ln)
    x = conv2d_same(x, P['head.kernel'], P['head.bias'])
    if is_A:
        x = P['tanh_scale'] * np.tanh(x)
    return x
