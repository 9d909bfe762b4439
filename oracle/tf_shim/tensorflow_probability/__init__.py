"""tfp.distributions.MultivariateNormalDiag for the reference's latent prior (M:1621-1623).  TEST INFRASTRUCTURE."""
import numpy as _np


class _MultivariateNormalDiag:
    def __init__(self, loc, scale_diag):
        self.loc = _np.asarray(loc, dtype=_np.float64)
        self.scale = _np.asarray(scale_diag, dtype=_np.float64)

    def log_prob(self, x):
        """log N(x; loc, diag(scale^2)) over the LAST axis (event shape [k]); leading axes are batch."""
        x = _np.asarray(x, dtype=_np.float64)
        z = (x - self.loc) / self.scale
        k = self.loc.shape[0]
        return -0.5 * _np.sum(z * z, axis=-1) - _np.sum(_np.log(self.scale)) - 0.5 * k * _np.log(2.0 * _np.pi)

    def sample(self, sample_shape=(), seed=None):
        shp = tuple(_np.atleast_1d(sample_shape).astype(int)) + self.loc.shape
        return self.loc + self.scale * _np.random.default_rng(seed).standard_normal(shp)


class distributions:
    MultivariateNormalDiag = _MultivariateNormalDiag
