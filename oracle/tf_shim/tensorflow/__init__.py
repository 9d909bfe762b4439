"""NumPy stand-in for the `tf.*` ops the reference's model files call (oracle/tf_shim/README.md).  TEST INFRASTRUCTURE.
Every op follows the documented TensorFlow semantics; tensors are NumPy arrays (fp64 by default so that the fixtures carry no
rounding of their own)."""
import numpy as _np

from . import keras  # noqa: F401  (tf.keras.layers.Reshape, tf.keras.initializers.Ones, ...)

float32 = _np.float64      # the fixtures are generated in fp64: `dtype=tf.float32` in the reference becomes fp64 here
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64

DTYPE = _np.float64


def _a(x, dtype=None):
    if isinstance(x, keras.layers.KTensor):
        raise TypeError("raw tf op on a symbolic Keras tensor: not needed by the reference, not implemented in the shim")
    return _np.asarray(x, dtype=dtype)


def function(fn=None, **_kw):
    """@tf.function / @tf.function(...): eager execution."""
    if fn is None:
        return lambda f: f
    return fn


def constant(value, dtype=None):
    return _np.asarray(value, dtype=dtype)


def cast(x, dtype):
    return _np.asarray(x).astype(dtype)


def range(start, limit=None, delta=1, dtype=None, **kw):   # noqa: A001
    if "start" in kw:
        start = kw["start"]
    if limit is None:
        start, limit = 0, start
    out = _np.arange(int(start), int(limit), int(delta))
    return out.astype(dtype) if dtype is not None else out.astype(_np.int32)


def shape(x):
    return _np.asarray(_np.shape(x), dtype=_np.int32)


def ones(shape, dtype=None):       # noqa: A002
    shp = tuple(int(s) for s in _np.atleast_1d(_np.asarray(shape)))
    return _np.ones(shp, dtype=dtype or DTYPE)


def zeros(shape, dtype=None):      # noqa: A002
    shp = tuple(int(s) for s in _np.atleast_1d(_np.asarray(shape)))
    return _np.zeros(shp, dtype=dtype or DTYPE)


def stack(values, axis=0):
    return _np.stack([_np.asarray(v) for v in values], axis=axis)


def concat(values, axis):
    return _np.concatenate([_a(v) for v in values], axis=int(axis))


def tile(x, multiples):
    return _np.tile(_a(x), tuple(int(m) for m in _np.asarray(multiples)))


def expand_dims(x, axis):
    return _np.expand_dims(_a(x), int(axis))


def transpose(x, perm=None):
    return _np.transpose(_a(x), None if perm is None else tuple(int(p) for p in perm))


def reshape(x, shape):             # noqa: A002
    return _np.reshape(_a(x), tuple(int(s) for s in shape))


def squeeze(x, axis=None):
    return _np.squeeze(_a(x), axis=_axes(axis))


def repeat(x, repeats, axis=None):
    return _np.repeat(_a(x), repeats, axis=axis)


class _ListDataset:
    """tf.data.Dataset stand-in: an eager list of elements with map() (what preprocess_dataset_* use, F:174-279)."""

    def __init__(self, elements):
        self.elements = list(elements)

    def map(self, fn, num_parallel_calls=None):   # noqa: A003
        return _ListDataset(fn(_np.array(e, dtype=DTYPE)) for e in self.elements)

    def __iter__(self):
        return iter(self.elements)


class _Data:
    AUTOTUNE = -1

    class Dataset:
        @staticmethod
        def from_tensor_slices(x):
            return _ListDataset(_np.asarray(x))


data = _Data()


def ensure_shape(x, shape):        # noqa: A002
    got = _np.shape(x)
    want = list(shape)
    if len(got) != len(want) or any(w is not None and int(w) != g for g, w in zip(got, want)):
        raise ValueError(f"Shape of tensor {got} is not compatible with expected shape {tuple(want)}.")
    return x


def scatter_nd(indices, updates, shape):   # noqa: A002
    """tf.scatter_nd: zeros(shape) with `updates[i]` ADDED at index `indices[i]` (indices of shape (n, 1) address the first
    axis, which is the only form the reference uses)."""
    indices = _np.asarray(indices)
    updates = _a(updates)
    shp = tuple(int(s) for s in _np.asarray(shape))
    assert indices.ndim == 2 and indices.shape[1] == 1, "the shim implements first-axis scatter only"
    out = _np.zeros(shp, dtype=updates.dtype)
    _np.add.at(out, indices[:, 0], updates)
    return out


def einsum(eq, *ops):
    return _np.einsum(eq, *[_a(o) for o in ops])


def exp(x):
    return _np.exp(_a(x))


def abs(x):    # noqa: A001
    return _np.abs(_a(x))


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


def reduce_sum(x, axis=None, keepdims=False):
    return _np.sum(_a(x), axis=_axes(axis), keepdims=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    return _np.mean(_a(x), axis=_axes(axis), keepdims=keepdims)


class _Math:
    reduce_sum = staticmethod(reduce_sum)
    reduce_mean = staticmethod(reduce_mean)
    exp = staticmethod(exp)
    abs = staticmethod(abs)

    @staticmethod
    def multiply(x, y):
        return _a(x) * _a(y)

    @staticmethod
    def scalar_mul(scalar, x):
        return _a(scalar) * _a(x)

    @staticmethod
    def reciprocal(x):
        return 1.0 / _a(x)

    @staticmethod
    def log(x):
        return _np.log(_a(x))

    @staticmethod
    def abs(x):      # noqa: A003
        return _np.abs(_a(x))

    @staticmethod
    def floor(x):
        return _np.floor(x)

    @staticmethod
    def ceil(x):
        return _np.ceil(x)


math = _Math()


class _NN:
    @staticmethod
    def space_to_depth(x, block_size):
        """NHWC: out[b, i, j, (dy * bs + dx) * C + c] = in[b, i * bs + dy, j * bs + dx, c]."""
        x = _a(x)
        bs = int(block_size)
        B, H, W, C = x.shape
        x = x.reshape(B, H // bs, bs, W // bs, bs, C).transpose(0, 1, 3, 2, 4, 5)
        return _np.ascontiguousarray(x.reshape(B, H // bs, W // bs, bs * bs * C))

    @staticmethod
    def depth_to_space(x, block_size):
        x = _a(x)
        bs = int(block_size)
        B, h, w, C4 = x.shape
        C = C4 // (bs * bs)
        x = x.reshape(B, h, w, bs, bs, C).transpose(0, 1, 3, 2, 4, 5)
        return _np.ascontiguousarray(x.reshape(B, h * bs, w * bs, C))


nn = _NN()


class _LinearOperatorFullMatrix:
    def __init__(self, matrix):
        self.m = _a(matrix, DTYPE)

    def matvec(self, x):
        return _np.einsum('ij,...j->...i', self.m, _a(x, DTYPE))


class _LinearOperatorDiag:
    def __init__(self, diag):
        self.d = _a(diag, DTYPE)

    def matvec(self, x):
        return self.d * _a(x, DTYPE)


class _Linalg:
    """tf.linalg: what TOYcINN_make_model.py calls (T:379-436)."""
    LinearOperatorFullMatrix = _LinearOperatorFullMatrix
    LinearOperatorDiag = _LinearOperatorDiag

    @staticmethod
    def inv(op):
        assert isinstance(op, _LinearOperatorDiag)
        return _LinearOperatorDiag(1.0 / op.d)

    @staticmethod
    def det(op):
        assert isinstance(op, _LinearOperatorDiag)
        return _np.prod(op.d, axis=-1)


linalg = _Linalg()


class GradientTape:      # cFlow.train_step is not exercised under the shim
    def __enter__(self):
        raise NotImplementedError("tf.GradientTape: the shim runs the forward / inverse / loss paths only")

    def __exit__(self, *a):
        return False
