"""Initializers: only what cFlow's signature and tanh_scaling_layer need.  The fixtures assign every weight explicitly."""
import numpy as _np


class Ones:
    def __call__(self, shape, dtype=None):
        return _np.ones(shape, dtype=dtype or _np.float64)


class Zeros:
    def __call__(self, shape, dtype=None):
        return _np.zeros(shape, dtype=dtype or _np.float64)


class Orthogonal:
    """Stand-in: a seeded Gaussian scaled by `gain` (same shapes; the values are overwritten before anything is computed)."""

    def __init__(self, gain=1.0, seed=None):
        self.gain, self.rng = gain, _np.random.default_rng(seed)

    def __call__(self, shape, dtype=None):
        fan_in = max(1, int(_np.prod(shape[:-1])))
        return (self.gain * self.rng.standard_normal(shape) / _np.sqrt(fan_in)).astype(dtype or _np.float64)


class GlorotUniform(Orthogonal):
    pass


def get(identifier):
    if isinstance(identifier, str) and identifier.lower() in ('ones', 'zeros'):
        return Ones() if identifier.lower() == 'ones' else Zeros()
    if identifier is None or isinstance(identifier, str):
        return GlorotUniform()
    return identifier
