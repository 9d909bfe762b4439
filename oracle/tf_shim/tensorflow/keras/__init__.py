"""tf.keras subset for the reference's model files (oracle/tf_shim/README.md).  TEST INFRASTRUCTURE."""
from . import backend, initializers, layers, metrics, regularizers  # noqa: F401
from .layers import Model  # noqa: F401
