"""Drop-in mirror of the reference's `TOYcINN_make_model.py` (T) on the B200-native path.

`coupling_layer(u1_size, u2_size, intermediate_dims, num_layers)` (T:29-97) and
`cINN_affine(io_shape, x_d, num_coupling_layers, intermediate_dims, num_layers, init, mask_indices=None)`
(T:105-506) keep their signatures and attributes; `model(u, direction)` returns `(v, log_detJ)` in both
directions with the TOY sign convention (-1: xy -> zy, +1: zy -> xy; Q2).  The whole flow runs as ONE
sm_100a kernel (csrc/toy_kernels.cu).
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, Borrowed, stream_ptr, require_cuda, int_array
from .conv_cINN_make_model import _Mean, _StandardNormalDiag, _default_device


class coupling_layer:
    """Dense s/t networks of one toy coupling layer (T:29-97): Dense(u1->I)+LReLU, num_layers x
    [Dense(I->I)+LReLU], Dense(I->u2); A ends in tanh.  Keras Dense default glorot-uniform kernels and
    zero biases (the model-level `init` argument is stored but unused by the reference, T:138)."""

    def __init__(self, u1_size, u2_size, intermediate_dims, num_layers, rng=None):
        self.u1_size, self.u2_size = int(u1_size), int(u2_size)
        self.intermediate_dims, self.num_layers = int(intermediate_dims), int(num_layers)
        rng = rng if rng is not None else np.random.default_rng()
        dims = [self.u1_size] + [self.intermediate_dims] * (self.num_layers + 1) + [self.u2_size]
        self._w = {}
        for net in ("b", "A"):                                # creation order in the reference (T:52-93)
            layers = []
            for a, b in zip(dims[:-1], dims[1:]):
                lim = math.sqrt(6.0 / (a + b))
                layers.append([rng.uniform(-lim, lim, (a, b)).astype(np.float32), np.zeros(b, np.float32)])
            self._w[net] = layers
        self._owner = None

    def get_weights(self):
        """Flat list [W, b, W, b, ...] for net b then net A (Keras layer-creation order, T:52-93)."""
        if self._owner is not None:
            self._owner.sync_layer_weights()
        out = []
        for net in ("b", "A"):
            for W, b in self._w[net]:
                out += [W.copy(), b.copy()]
        return out

    def set_weights(self, weights):
        n = len(self._w["b"])
        assert len(weights) == 4 * n, f"expected {4 * n} arrays, got {len(weights)}"
        if self._owner is not None:
            self._owner.sync_layer_weights()       # keep what a previous train_step wrote into the other layers
        k = 0
        for net in ("b", "A"):
            for i in range(n):
                W, b = np.asarray(weights[k], np.float32), np.asarray(weights[k + 1], np.float32)
                assert W.shape == self._w[net][i][0].shape and b.shape == self._w[net][i][1].shape
                self._w[net][i] = [W.copy(), b.copy()]
                k += 2
        if self._owner is not None:
            self._owner._dirty = True

    def __call__(self, u_1):
        raise NotImplementedError("the toy s/t MLPs only run fused inside cINN_affine's flow kernel")


class cINN_affine:
    def __init__(self, io_shape, x_d, num_coupling_layers, intermediate_dims, num_layers, init,
                 mask_indices=None, device=None, seed=None):
        self.io_shape = io_shape
        self.x_d = x_d
        self.num_coupling_layers = num_coupling_layers
        self.intermediate_dims = intermediate_dims
        self.num_layers = num_layers
        self.init = init
        self.lambda_y = 100
        self.device = torch.device(device) if device is not None else _default_device()
        if io_shape != 3:
            raise NotImplementedError("the reference's masks are written for 3-dimensional xy (T:154-166)")
        self.distribution = _StandardNormalDiag(self.x_d, self.device)       # T:141-143
        self.mask_dict_1 = {0: np.array([0]), 1: np.array([1]), 2: np.array([2]),
                            3: np.array([0, 1]), 4: np.array([0, 2]), 5: np.array([1, 2])}
        self.mask_dict_2 = {0: np.array([1, 2]), 1: np.array([0, 2]), 2: np.array([0, 1]),
                            3: np.array([2]), 4: np.array([1]), 5: np.array([0])}
        identity = np.identity(io_shape, dtype=np.float32)
        self.masks_1, self.masks_2 = {}, {}
        if mask_indices is not None and len(mask_indices):
            self.mask_indices = mask_indices
        else:
            self.mask_indices = np.arange(num_coupling_layers, dtype=np.int32)
        for i in self.mask_indices:                           # T:176-178
            self.masks_1[i] = identity[self.mask_dict_1[i % 6]]
            self.masks_2[i] = identity[self.mask_dict_2[i % 6]]
        self.dims_u1 = np.zeros_like(self.mask_indices, dtype=np.int32)
        self.dims_u2 = np.zeros_like(self.mask_indices, dtype=np.int32)
        for i in self.mask_indices:
            self.dims_u1[i] = np.sum(self.masks_1[i], dtype=np.int32)
            self.dims_u2[i] = np.sum(self.masks_2[i], dtype=np.int32)
        rng = np.random.default_rng(seed)
        self.coupling_layers_list = [coupling_layer(int(self.dims_u1[i]), int(self.dims_u2[i]),
                                                    intermediate_dims, num_layers, rng)
                                     for i in range(num_coupling_layers)]
        for cl in self.coupling_layers_list:
            cl._owner = self
        if mask_indices is None or not len(mask_indices):     # T:206-217: shuffle within groups of six
            shuffler = np.array([self.mask_indices[6 * i:6 * (i + 1)] for i in range(num_coupling_layers // 6)])
            for i in range(num_coupling_layers // 6):
                rng.shuffle(shuffler[i])
            self.mask_indices = shuffler.flatten()
        self.loss_tracker = _Mean('loss')
        self.z_loss_tracker = _Mean('z_loss')
        self.y_loss_tracker = _Mean('y_loss')
        self.detJ_loss_tracker = _Mean('detJ_loss')
        self.optimizer = None
        self.params = None
        self._dirty = True
        self.last_per_sample = None

    @property
    def metrics(self):
        return [self.loss_tracker, self.z_loss_tracker, self.y_loss_tracker, self.detJ_loss_tracker]

    def compile(self, optimizer=None):
        self.optimizer = optimizer

    # -- flat device parameter buffer in the layout of csrc/toy_kernels.cu ----------------------------
    def _pack(self):
        I, L, n = self.intermediate_dims, self.num_layers, self.num_coupling_layers
        total = int(lib.cnf_toy_param_count(n, I, L))
        if total <= 0:
            raise NotImplementedError("unsupported toy configuration")
        net_sz = int(lib.cnf_toy_layer_offset(1, I, L)) // 2
        flat = np.zeros(total, np.float32)
        for j, cl in enumerate(self.coupling_layers_list):
            for k, net in enumerate(("A", "b")):
                base = int(lib.cnf_toy_layer_offset(j, I, L)) + k * net_sz
                layers = cl._w[net]
                W0, b0 = layers[0]
                flat[base: base + W0.shape[0] * I] = W0.reshape(-1)         # [2][I], rows >= dim(u1) unused
                flat[base + 2 * I: base + 3 * I] = b0
                p = base + 3 * I
                for W, b in layers[1:-1]:
                    flat[p: p + I * I] = W.reshape(-1)
                    flat[p + I * I: p + I * I + I] = b
                    p += I * I + I
                WL, bL = layers[-1]
                wl = np.zeros((I, 2), np.float32)
                wl[:, :WL.shape[1]] = WL
                flat[p: p + 2 * I] = wl.reshape(-1)
                flat[p + 2 * I: p + 2 * I + bL.shape[0]] = bL
        self.params = torch.from_numpy(flat).to(self.device)
        self._dirty = False

    def _order(self):
        return int_array(int(i) for i in self.mask_indices)

    # -- cINN_affine.call (T:248-402) ------------------------------------------------------------------
    def call(self, u, direction=-1):
        u = require_cuda(u, "u")
        if u.dim() != 2 or u.shape[1] != 3:
            raise ValueError(f"u: expected shape [B, 3], got {tuple(u.shape)}")
        if self._dirty or self.params is None:
            self._pack()
        B = u.shape[0]
        v = torch.empty_like(u)
        ld = torch.empty(B, dtype=torch.float32, device=u.device)
        br = Borrowed()
        check(lib.cnf_toy_call(br(u), br(self.params), self._order(), self.num_coupling_layers,
                               self.intermediate_dims, self.num_layers, int(direction), br(v), br(ld),
                               stream_ptr()))
        return v, (ld if direction == -1 else 0)              # T:402 (log_detJ stays the int 0 for +1)

    __call__ = call

    # -- log_loss (T:404-451) ------------------------------------------------------------------------------
    def log_loss(self, xy):
        xy = require_cuda(xy, "xy")
        if xy.dim() != 2 or xy.shape[1] != 3:
            raise ValueError(f"xy: expected shape [B, 3], got {tuple(xy.shape)}")
        if self._dirty or self.params is None:
            self._pack()
        B = xy.shape[0]
        Bp = (B + 3) & ~3
        zy = torch.empty_like(xy)
        pers = torch.empty((3, Bp), dtype=torch.float32, device=xy.device)
        loss4 = torch.empty(4, dtype=torch.float32, device=xy.device)
        ll_z, ll_y, ld = pers[0, :B], pers[1, :B], pers[2, :B]
        br = Borrowed()
        check(lib.cnf_toy_log_loss(br(xy), br(self.params), self._order(), self.num_coupling_layers,
                                   self.intermediate_dims, self.num_layers, int(self.x_d), float(self.lambda_y),
                                   br(zy), br(ll_z), br(ll_y), br(ld), br(loss4), stream_ptr()))
        self.last_per_sample = {'ll_z': ll_z, 'll_y': ll_y, 'logdet': ld, 'zy': zy}
        return loss4[0], loss4[1], loss4[2], loss4[3]

    def _update_trackers(self, four):
        vals = [float(v) for v in torch.stack(list(four)).cpu()]
        for tr, v in zip(self.metrics, vals):
            tr.update_state(v)
        return {'loss': self.loss_tracker.result(), 'z_loss': self.z_loss_tracker.result(),
                'y_loss': self.y_loss_tracker.result(), 'detJ_loss': self.detJ_loss_tracker.result()}

    def test_step(self, xy):
        return self._update_trackers(self.log_loss(xy))

    # -- training (T:453-482) -----------------------------------------------------------------------------------
    def loss_and_grad(self, xy):
        """log_loss(xy) plus d loss / d params as ONE flat tensor with the layout of self.params (hand-written backward
        kernel, csrc/toy_kernels.cu: no stored activations, layer inputs recovered with the inverse law).
        Returns ((loss, z_loss, y_loss, detJ_loss), grads)."""
        xy = require_cuda(xy, "xy")
        if xy.dim() != 2 or xy.shape[1] != 3:
            raise ValueError(f"xy: expected shape [B, 3], got {tuple(xy.shape)}")
        if self._dirty or self.params is None:
            self._pack()
        B = xy.shape[0]
        Bp = (B + 3) & ~3
        zy = torch.empty_like(xy)
        pers = torch.empty((3, Bp), dtype=torch.float32, device=xy.device)
        loss4 = torch.empty(4, dtype=torch.float32, device=xy.device)
        ll_z, ll_y, ld = pers[0, :B], pers[1, :B], pers[2, :B]
        if getattr(self, '_grads', None) is None or self._grads.shape != self.params.shape:
            self._grads = torch.empty_like(self.params)
        br = Borrowed()
        check(lib.cnf_toy_loss_and_grad(br(xy), br(self.params), self._order(), self.num_coupling_layers,
                                        self.intermediate_dims, self.num_layers, int(self.x_d), float(self.lambda_y),
                                        br(self._grads), br(zy), br(ll_z), br(ll_y), br(ld), br(loss4), stream_ptr()))
        self.last_per_sample = {'ll_z': ll_z, 'll_y': ll_y, 'logdet': ld, 'zy': zy}
        return (loss4[0], loss4[1], loss4[2], loss4[3]), self._grads

    def _unpack(self, flat=None):
        """flat device buffer (self.params, or a gradient buffer of the same layout) -> per-layer Keras-shaped lists
        [{'A': [(W, b), ...], 'b': [...]}, ...]; with flat=None the layers' own weights are refreshed from self.params."""
        I, L = self.intermediate_dims, self.num_layers
        host = (self.params if flat is None else flat).detach().cpu().numpy()
        net_sz = int(lib.cnf_toy_layer_offset(1, I, L)) // 2
        out = []
        for j, cl in enumerate(self.coupling_layers_list):
            entry = {}
            for k, net in enumerate(("A", "b")):
                base = int(lib.cnf_toy_layer_offset(j, I, L)) + k * net_sz
                d1, d2 = cl.u1_size, cl.u2_size
                layers = [(host[base: base + d1 * I].reshape(d1, I).copy(), host[base + 2 * I: base + 3 * I].copy())]
                p = base + 3 * I
                for _ in range(L):
                    layers.append((host[p: p + I * I].reshape(I, I).copy(), host[p + I * I: p + I * I + I].copy()))
                    p += I * I + I
                layers.append((host[p: p + 2 * I].reshape(I, 2)[:, :d2].copy(), host[p + 2 * I: p + 2 * I + d2].copy()))
                entry[net] = layers
            out.append(entry)
        if flat is None:
            for cl, entry in zip(self.coupling_layers_list, out):
                for net in ("A", "b"):
                    cl._w[net] = [[W, b] for W, b in entry[net]]
        return out

    def grad_views(self):
        """the last gradient as [{'A': [(dW, db), ...], 'b': [...]}, ...] (same nesting as the oracle's weights)"""
        return self._unpack(self._grads)

    def train_step(self, xy):
        """cINN_affine.train_step (T:453-482): gradients of log_loss, optimizer.apply_gradients, metric trackers."""
        if self.optimizer is None:
            raise RuntimeError("train_step: call model.compile(optimizer=Adam(...)) first")
        four, grads = self.loss_and_grad(xy)
        self.optimizer.apply_gradients(self.params, grads)
        self._layers_stale = True          # coupling_layer.get_weights() refreshes from the device buffer on demand
        return self._update_trackers(four)

    def sync_layer_weights(self):
        """copy the trained flat buffer back into coupling_layers_list[i] (get_weights of the layers, Y:228-235)"""
        if getattr(self, '_layers_stale', False):
            self._unpack()
            self._layers_stale = False
