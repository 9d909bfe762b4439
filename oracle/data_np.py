"""TEST INFRASTRUCTURE (oracle): NumPy restatement of the data helpers of conv_cINN_base_functions.py (F) that sit on
either side of the flow.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

  down / up                 F:74-164
  preprocess_class_logits   F:174-231  (the LOGITS=True branch)
  preprocess_SR             F:233-279
  de_logitify               F:287-318
  instance_noise            F:635-653  (noise supplied by the caller, or from the Philox restatement below)

TensorFlow's random stream is not reproducible outside TensorFlow, so the noise generator is pinned differently: the
product uses counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
SC'11) + Box-Muller; `philox4x32_10` restates the published algorithm and is checked against the Random123
known-answer vectors in tests/test_oracle_data.py.
"""
import numpy as np


def down(img):
    """F:74-127: reshape to (B, H/2, 2, W/2, 2, D) and reduce_mean over axes (2, 4); crops odd trailing rows/cols."""
    img = np.asarray(img)
    batch = img.ndim == 4
    if not batch:
        img = img[None]
    B, M, N, D = img.shape
    MK, NL = M // 2, N // 2
    x = img[:, :MK * 2, :NL * 2, :].reshape(B, MK, 2, NL, 2, D)
    # ((a + b) + (c + d)) / 4: pairwise over the column pair, then over the row pair, in the array's own dtype
    s = (x[:, :, 0, :, 0, :] + x[:, :, 0, :, 1, :]) + (x[:, :, 1, :, 0, :] + x[:, :, 1, :, 1, :])
    out = (s * img.dtype.type(0.25)).astype(img.dtype)
    return out if batch else out[0]


def up(img):
    """F:129-164: tf.repeat(repeats=2) along H then W."""
    img = np.asarray(img)
    batch = img.ndim == 4
    if not batch:
        img = img[None]
    out = np.repeat(np.repeat(img, 2, axis=1), 2, axis=2)
    return out if batch else out[0]


def _nest(f, x, n):
    for _ in range(n):
        x = f(x)
    return x


def preprocess_SR(hires, model_type=None, RESIDUAL=True, levels=None):
    """F:233-279.  'SR4,2': x = down(h), y = up(down(down(h))); 'SR2,1': x = h, y = up(down(h)); levels=(lx, ly)
    generalises: x = down^lx(h), y = up^(ly-lx)(down^ly(h))."""
    if levels is None:
        levels = {'SR4,2': (1, 2), 'SR2,1': (0, 1)}[model_type]
    lx, ly = levels
    x = _nest(down, hires, lx)
    y = _nest(up, _nest(down, hires, ly), ly - lx)
    if RESIDUAL:
        x = x - y
    return np.concatenate((x, y), axis=-1)


def _logit(x):
    return np.log(x / (1 - x))


def preprocess_class_logits(x, a=0.01):
    """F:199-231: x -> (logit(a + (1-a) b x) - logit(a)) / (logit(1-a) - logit(a)), b = (1-2a)/(1-a)."""
    x = np.asarray(x, dtype=np.float64)
    b = (1 - 2 * a) / (1 - a)
    mn, mx = _logit(a), _logit(1 - a)
    return (_logit(a + (1 - a) * b * x) - mn) / (mx - mn)


def de_logitify(x, a=0.01):
    """F:287-318."""
    x = np.asarray(x, dtype=np.float64)
    mn, mx = _logit(a), _logit(1 - a)
    b = (1 - 2 * a) / (1 - a)
    x = x * (mx - mn) + mn
    return (1 / (1 + np.exp(-x)) - a) / (b * (1 - a))


def instance_noise(x, alpha, noise):
    """F:635-653 with the N(0,1) draw supplied: alpha x + (1 - alpha) noise."""
    return alpha * np.asarray(x, dtype=np.float64) + (1 - alpha) * np.asarray(noise, dtype=np.float64)


# ---- Philox4x32-10 ---------------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32 array [..., 4]; key: (k0, k1).  Returns uint32 [..., 4].  Ten rounds of
    (c0, c1, c2, c3) <- (hi(M1 c2) ^ c1 ^ k0, lo(M1 c2), hi(M0 c0) ^ c3 ^ k1, lo(M0 c0)); key += (W0, W1)."""
    c = np.asarray(ctr, dtype=np.uint64).copy()
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c[..., 0]
        p1 = _M1 * c[..., 2]
        n0 = (p1 >> np.uint64(32)) ^ c[..., 1] ^ np.uint64(k0)
        n1 = p1 & _MASK
        n2 = (p0 >> np.uint64(32)) ^ c[..., 3] ^ np.uint64(k1)
        n3 = p0 & _MASK
        c = np.stack([n0, n1, n2, n3], axis=-1)
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def philox_normal(n, seed, offset=0):
    """The N(0,1) stream of cnf_instance_noise in float64: elements 4q..4q+3 come from Philox counter (q + offset, 0)
    under key (seed lo, seed hi); Box-Muller on (u0, u1) and (u2, u3) with u_a = (k + 1) 2^-32 and u_b = k 2^-32:
    z0 = r cos(2 pi u_b), z1 = r sin(2 pi u_b), r = sqrt(-2 ln u_a)."""
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64) + np.uint64(offset)
    ctr = np.zeros((nq, 4), dtype=np.uint64)
    ctr[:, 0] = q & _MASK
    ctr[:, 1] = q >> np.uint64(32)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.float64)
    out = np.empty((nq, 4))
    for j in (0, 2):
        ua = (r[:, j] + 1.0) * 2.0 ** -32
        ub = r[:, j + 1] * 2.0 ** -32
        rad = np.sqrt(-2.0 * np.log(ua))
        out[:, j] = rad * np.cos(2 * np.pi * ub)
        out[:, j + 1] = rad * np.sin(2 * np.pi * ub)
    return out.reshape(-1)[:n]
