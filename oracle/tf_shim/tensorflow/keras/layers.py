"""tf.keras functional-API subset (oracle/tf_shim/README.md).  TEST INFRASTRUCTURE.

A symbolic tensor (`KTensor`) carries a batch-1 dummy value, so static shapes come from actually running the layer once; the
node graph is re-executed layer by layer on every `Model.__call__`, as tf.keras >= 2.4 does.  Layer and variable names follow
Keras: snake-cased class name plus a per-name counter (`conv2d`, `conv2d_1`, ...), `<layer>/kernel:0`, `<layer>/bias:0`,
`<layer>/gamma:0`, `<layer>/beta:0`, `<layer>/Variable:0` for an unnamed `add_weight`.
"""
import re

import numpy as _np
import torch as _torch
import torch.nn.functional as _F

from . import initializers as _init

#: True: a Lambda's Python function runs again on every model call (tf.keras >= 2.4: late-bound closure variables take their
#: FINAL value).  False: the slice it took while the graph was built is frozen (TF 1.x graph-mode Keras: run once).
LAMBDA_REPLAY = True

_UIDS = {}
_VARIABLES = []          # every variable of the process in creation order


def reset_state():
    _UIDS.clear()
    del _VARIABLES[:]


def all_variables():
    return list(_VARIABLES)


def _snake(name):
    s = re.sub('(.)([A-Z][a-z0-9]+)', r'\1_\2', name)
    return re.sub('([a-z])([A-Z])', r'\1_\2', s).lower()


def _unique(base):
    k = _UIDS.get(base, 0)
    _UIDS[base] = k + 1
    return base if k == 0 else f"{base}_{k}"


class Variable:
    def __init__(self, name, value):
        self.name, self.value = name, _np.array(value, dtype=_np.float64)
        _VARIABLES.append(self)

    @property
    def shape(self):
        return self.value.shape

    def numpy(self):
        return self.value

    def assign(self, v):
        v = _np.asarray(v, dtype=_np.float64)
        assert v.shape == self.value.shape, (self.name, v.shape, self.value.shape)
        self.value[...] = v

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)


class KTensor:
    """symbolic tensor: `dummy` is the value a batch of ONE zero sample produces; `node` = (layer, inputs)"""

    def __init__(self, dummy, node=None):
        self.dummy, self.node = _np.asarray(dummy), node

    @property
    def shape(self):
        return (None,) + tuple(int(s) for s in self.dummy.shape[1:])

    dtype = _np.float64


def _is_sym(x):
    if isinstance(x, (list, tuple)):
        return any(_is_sym(v) for v in x)
    return isinstance(x, KTensor)


def _dummies(x):
    if isinstance(x, (list, tuple)):
        return [_dummies(v) for v in x]
    return x.dummy if isinstance(x, KTensor) else x


class Layer:
    _base_name = None

    def __init__(self, name=None, **kwargs):
        self.name = name or _unique(self._base_name or _snake(type(self).__name__))
        self.built = False
        self._vars = []

    def add_weight(self, name=None, shape=(), initializer=None, trainable=True, dtype=None):
        initializer = _init.get(initializer) if not callable(initializer) else initializer
        v = Variable(f"{self.name}/{name or 'Variable'}:0", initializer(tuple(shape) if shape is not None else ()))
        self._vars.append(v)
        return v

    def build(self, input_shape):
        pass

    def call(self, inputs):
        raise NotImplementedError(str(type(self)))

    def get_config(self):
        return {'name': self.name}

    @property
    def weights(self):
        return list(self._vars)

    def _run(self, inputs, building=False):
        if not self.built:
            shp = [(None,) + tuple(_np.shape(v)[1:]) for v in inputs] if isinstance(inputs, (list, tuple)) \
                else (None,) + tuple(_np.shape(inputs)[1:])
            self.build(shp)
            self.built = True
        return self.call(inputs)

    def __call__(self, inputs, *args, **kwargs):
        if _is_sym(inputs):
            return KTensor(self._run(_dummies(inputs), building=True), node=(self, inputs))
        return self._run(inputs)


def Input(shape, dtype=None, **kw):     # noqa: N802
    shape = (shape,) if _np.isscalar(shape) else tuple(shape)       # Input(shape=3) is accepted by Keras (T:48)
    return KTensor(_np.zeros((1,) + tuple(int(s) for s in shape), dtype=_np.float64), node=None)


class Model(Layer):
    """Functional (`Model(inputs=, outputs=)`) or subclassed (`class cFlow(Model)`: __call__ -> self.call)."""

    def __init__(self, inputs=None, outputs=None, name=None, **kwargs):
        super().__init__(name=name)
        self.inputs, self.outputs = inputs, outputs

    def __call__(self, *args, **kwargs):
        if self.outputs is None:
            return self.call(*args, **kwargs)
        x = args[0]
        memo = {id(self.inputs): _np.asarray(x, dtype=_np.float64)}

        def value(t):
            if id(t) in memo:
                return memo[id(t)]
            layer, ins = t.node
            vals = [value(v) for v in ins] if isinstance(ins, (list, tuple)) else value(ins)
            out = layer._run(vals)
            memo[id(t)] = out
            return out
        if isinstance(self.outputs, (list, tuple)):
            return [value(t) for t in self.outputs]
        return value(self.outputs)


class Conv2D(Layer):
    _base_name = 'conv2d'

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', dilation_rate=(1, 1), use_bias=True,
                 kernel_initializer='glorot_uniform', **kwargs):
        super().__init__(**kwargs)
        pair = lambda v: (int(v), int(v)) if _np.isscalar(v) else tuple(int(s) for s in v)  # noqa: E731
        self.filters, self.ks, self.strides, self.padding = int(filters), pair(kernel_size), pair(strides), padding
        self.dil, self.use_bias, self.kinit = pair(dilation_rate), use_bias, kernel_initializer

    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight('kernel', self.ks + (cin, self.filters), self.kinit)
        self.bias = self.add_weight('bias', (self.filters,), _init.Zeros()) if self.use_bias else None

    def call(self, x):
        assert self.strides == (1, 1), "the reference only uses stride 1"
        xt = _torch.from_numpy(_np.ascontiguousarray(_np.asarray(x, dtype=_np.float64))).permute(0, 3, 1, 2)
        w = _torch.from_numpy(self.kernel.value).permute(3, 2, 0, 1).contiguous()      # HWIO -> OIHW (cross-correlation)
        if self.padding == 'same':        # stride 1: total = d (k - 1), the smaller half in front (TensorFlow's rule)
            th, tw = self.dil[0] * (self.ks[0] - 1), self.dil[1] * (self.ks[1] - 1)
            xt = _F.pad(xt, (tw // 2, tw - tw // 2, th // 2, th - th // 2))
        y = _F.conv2d(xt, w, None if self.bias is None else _torch.from_numpy(self.bias.value), dilation=self.dil)
        return y.permute(0, 2, 3, 1).contiguous().numpy()


Convolution2D = Conv2D


class Dense(Layer):
    _base_name = 'dense'

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer='glorot_uniform', kernel_regularizer=None,
                 **kwargs):
        super().__init__(**kwargs)
        assert activation is None
        self.units, self.use_bias, self.kinit = int(units), use_bias, kernel_initializer

    def build(self, input_shape):
        self.kernel = self.add_weight('kernel', (int(input_shape[-1]), self.units), self.kinit)
        self.bias = self.add_weight('bias', (self.units,), _init.Zeros()) if self.use_bias else None

    def call(self, x):
        y = _np.asarray(x, dtype=_np.float64) @ self.kernel.value
        return y + self.bias.value if self.bias is not None else y


class LayerNormalization(Layer):
    _base_name = 'layer_normalization'

    def __init__(self, axis=-1, epsilon=1e-3, center=True, scale=True, **kwargs):
        super().__init__(**kwargs)
        self.axis, self.eps = axis, epsilon

    def build(self, input_shape):
        n = int(input_shape[self.axis])
        self.gamma = self.add_weight('gamma', (n,), _init.Ones())
        self.beta = self.add_weight('beta', (n,), _init.Zeros())

    def call(self, x):
        assert self.axis in (-1, _np.ndim(x) - 1)
        x = _np.asarray(x, dtype=_np.float64)
        mean = x.mean(axis=-1, keepdims=True)
        var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)        # biased, like tf.nn.moments
        return (x - mean) / _np.sqrt(var + self.eps) * self.gamma.value + self.beta.value


class LeakyReLU(Layer):
    _base_name = 'leaky_re_lu'

    def __init__(self, alpha=0.3, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    def call(self, x):
        x = _np.asarray(x)
        return _np.where(x >= 0, x, self.alpha * x)


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.fn = {'tanh': _np.tanh, 'linear': lambda v: v}[activation]

    def call(self, x):
        return self.fn(_np.asarray(x))


class Dropout(Layer):
    def __init__(self, rate, **kwargs):
        super().__init__(**kwargs)

    def call(self, x):
        return x          # inference


class Reshape(Layer):
    def __init__(self, target_shape, **kwargs):
        super().__init__(**kwargs)
        self.target = tuple(int(s) for s in target_shape)

    def call(self, x):
        x = _np.asarray(x)
        return x.reshape((x.shape[0],) + self.target)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, xs):
        return _np.concatenate([_np.asarray(v) for v in xs], axis=self.axis)


class Add(Layer):
    def call(self, xs):
        out = _np.asarray(xs[0])
        for v in xs[1:]:
            out = out + _np.asarray(v)
        return out


def concatenate(inputs, axis=-1, **kw):
    return Concatenate(axis=axis)(inputs)


def add(inputs, **kw):
    return Add()(inputs)


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super().__init__(**kwargs)
        self.function = function
        self._frozen = None

    def _run(self, inputs, building=False):
        if building:
            # what TF 1.x graph-mode Keras keeps: the ops of this ONE execution.  The reference's lambdas are channel slices,
            # which a probe of channel indices identifies exactly.
            x = _np.asarray(inputs)
            probe = _np.broadcast_to(_np.arange(x.shape[-1], dtype=_np.float64), x.shape)
            got = _np.asarray(self.function(probe))
            idx = got.reshape(-1, got.shape[-1])[0].astype(int)
            assert _np.array_equal(got, probe[..., idx]), "Lambda is not a last-axis selection: freeze it by hand"
            self._frozen = idx
            return self.function(inputs)
        if LAMBDA_REPLAY or self._frozen is None:
            return self.function(inputs)
        return _np.asarray(inputs)[..., self._frozen]
