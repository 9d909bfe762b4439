"""tf.keras.backend.int_shape: the static shape as a tuple (batch = None on symbolic tensors)."""
import numpy as _np


def int_shape(x):
    from .layers import KTensor
    if isinstance(x, KTensor):
        return x.shape
    return tuple(int(s) for s in _np.shape(x))
