"""Kernel initialisers with Keras semantics (host side, NumPy).

The reference passes `tf.keras.initializers.Orthogonal(gain=0.1)` as `init` (conv_cINN_make_model.py
M:1442, TOYcINN.py Y:100); Keras flattens a kernel of shape (kh, kw, cin, cout) to
(kh*kw*cin, cout), takes the Q of a QR decomposition of a normal matrix (sign-fixed by diag(R)),
and scales by `gain`.  Random streams differ from TensorFlow's; only the distribution matches.
"""
import numpy as np


class Orthogonal:
    def __init__(self, gain=1.0, seed=None):
        self.gain = gain
        self.rng = np.random.default_rng(seed)

    def __call__(self, shape, dtype=np.float32):
        shape = tuple(int(s) for s in shape)
        rows = int(np.prod(shape[:-1]))
        cols = int(shape[-1])
        a = self.rng.standard_normal((max(rows, cols), min(rows, cols)))
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diag(r))
        if rows < cols:
            q = q.T
        return (self.gain * q.reshape(shape)).astype(dtype)


class GlorotUniform:
    def __init__(self, seed=None):
        self.rng = np.random.default_rng(seed)

    def __call__(self, shape, dtype=np.float32):
        shape = tuple(int(s) for s in shape)
        receptive = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return self.rng.uniform(-lim, lim, shape).astype(dtype)


def get(init):
    if init is None:
        return Orthogonal(gain=0.1)
    if callable(init):
        return init
    if isinstance(init, str):
        key = init.lower()
        if key == "orthogonal":
            return Orthogonal(gain=1.0)
        if key in ("glorot_uniform", "glorotuniform"):
            return GlorotUniform()
    raise ValueError(f"unknown initializer {init!r}")
