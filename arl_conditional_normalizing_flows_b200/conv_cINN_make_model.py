"""Drop-in mirror of the reference's `conv_cINN_make_model.py` on the B200-native path.

Same class names, constructor arguments, methods and error behaviour as the reference
(/root/reference/conv_cINN_make_model.py, cited as M:line); the arithmetic runs in the hand-written
sm_100a kernels of libcnf.so through the C-ABI in include/cnf.h.  Tensors are torch CUDA fp32
NHWC-contiguous instead of tf.Tensor.  Weights live in one flat fp32 device buffer per model
(`cFlow.params`); every coupling layer holds a view of its slice and exposes Keras-shaped named
views (`get_weights` / `set_weights`).

Supersets of the reference surface (needed for per-sample parity checks):
`cFlow.last_logdet_per_sample`, `cFlow.last_per_sample`, `cFlow.bits_per_dim`.
"""
import math
from ctypes import byref, c_int, c_void_p

import numpy as np
import torch
import torch.distributed

from . import _lib, initializers
from ._lib import lib, check, Borrowed, stream_ptr, require_cuda, int_array, LAYER_GRADS_READY_FN
from .conv_cINN_base_functions import dilated_residual_block  # noqa: F401  (reference import, M:24)


def _default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


class _Mean:
    """keras.metrics.Mean stand-in for the four loss trackers (M:1692-1695)."""

    def __init__(self, name):
        self.name = name
        self.reset_state()

    def reset_state(self):
        self.total, self.count = 0.0, 0

    reset_states = reset_state

    def update_state(self, value):
        self.total += float(value)
        self.count += 1

    def result(self):
        return self.total / self.count if self.count else 0.0


class _StandardNormalDiag:
    """tfp MultivariateNormalDiag(loc=[0]*x_d, scale_diag=[1]*x_d) (M:1621-1623): sample / log_prob."""

    def __init__(self, x_d, device):
        self.x_d, self.device = x_d, device

    def sample(self, n, generator=None):
        shape = (n, self.x_d) if isinstance(n, int) else tuple(n) + (self.x_d,)
        return torch.randn(shape, device=self.device, dtype=torch.float32, generator=generator)

    def log_prob(self, z):
        return -0.5 * (z * z).sum(-1) - 0.5 * self.x_d * math.log(2.0 * math.pi)


###############################################################################

class Layer:
    """Bidirectional layer protocol (M:62-89)."""

    def __init__(self, **kwargs):
        self.name = kwargs.get("name")

    def build(self, input_shape=None):
        self.built = True

    def forward_and_Jacobian(self, u, sum_log_det_J, z):
        raise NotImplementedError(str(type(self)))

    def backward(self, v, z):
        raise NotImplementedError(str(type(self)))

    def get_config(self):
        return {"name": self.name}


class tanh_scaling_layer(Layer):
    """One trainable scalar multiplying its input (M:97-122).  Inside a coupling layer the scalar is
    the `tanh_scale` entry of net A and the multiply is fused into the head kernel's epilogue."""

    def __init__(self, w=None, **kwargs):
        super().__init__(**kwargs)
        self.w = w if w is not None else torch.ones((), dtype=torch.float32, device=_default_device())

    def call(self, inputs):
        return self.w * inputs

    __call__ = call


class squeeze_layer(Layer):
    """tf.nn.space_to_depth / depth_to_space with block 2 on u and on the running zy (M:130-217)."""

    def forward_and_Jacobian(self, u, sum_log_det_J, zy):
        assert u.shape[1] % 2 == 0 and u.shape[2] % 2 == 0, \
            'u must have spatial dimensions divisible by 2.'
        v = _space_to_depth(u)
        if zy is not None:
            zy = _space_to_depth(zy)
        return v, sum_log_det_J, zy

    def backward(self, v, zy):
        assert v.shape[3] % 4 == 0, 'v must have channel dimensions divisible by 4.'
        u = _depth_to_space(v)
        if zy is not None:
            zy = _depth_to_space(zy)
        return u, zy


class factor_out_zy_layer(Layer):
    """Moves half of the channels into / out of the running zy (M:219-329).  Pure slicing."""

    def __init__(self, num_prev_factors, **kwargs):
        super().__init__(**kwargs)
        self.num_prev_factors = num_prev_factors

    def get_config(self):
        config = super().get_config()
        config.update({'num_prev_factors': self.num_prev_factors})
        return config

    def forward_and_Jacobian(self, u, sum_log_det_J, zy):
        split = u.shape[3] // 2
        factored_zy = u[..., :split]
        v = u[..., split:].contiguous()
        zy = torch.cat([zy, factored_zy], dim=3) if zy is not None else factored_zy.contiguous()
        return v, sum_log_det_J, zy

    def backward(self, v, zy):
        if v is None:
            split = zy.shape[3] // (2 ** self.num_prev_factors)
        else:
            split = v.shape[3]
        reintegrated_v = zy[..., -split:]
        zy = zy[..., :-split].contiguous()
        assert reintegrated_v.shape[3] == split
        u = torch.cat([reintegrated_v, v], dim=3) if v is not None else reintegrated_v.contiguous()
        return u, zy


def _space_to_depth(x):
    x = require_cuda(x, "u")
    B, H, W, C = x.shape
    out = torch.empty((B, H // 2, W // 2, 4 * C), dtype=torch.float32, device=x.device)
    br = Borrowed()
    check(lib.cnf_space_to_depth(br(x), br(out), stream_ptr()))
    return out


def _depth_to_space(x):
    x = require_cuda(x, "v")
    B, h, w, C4 = x.shape
    out = torch.empty((B, 2 * h, 2 * w, C4 // 4), dtype=torch.float32, device=x.device)
    br = Borrowed()
    check(lib.cnf_depth_to_space(br(x), br(out), stream_ptr()))
    return out


###############################################################################

_ROLE = {0: "kernel", 1: "bias", 2: "gamma", 3: "beta", 4: "scale"}


class _NetHandle:
    """Callable stand-in for the Keras functional models `model_A` / `model_b` (M:1201-1204)."""

    def __init__(self, layer, which):
        self.layer, self.which = layer, which

    def __call__(self, u1_compressed):
        A, b = self.layer._nets(u1_compressed)
        return A if self.which == "A" else b

    def get_weights(self):
        return self.layer.get_weights()[self.which]


class coupling_layer(Layer):
    """RealNVP coupling layer on (xy -> zy) with compressed masks and ResNeXt s/t nets (M:331-1394)."""

    def __init__(self, in_shape, which_mask, num_res_blocks, cardinality, num_kernels, kernel_size, init,
                 LAYER_NORM=False, which_dilations=[1, 2, 4], _handle=None, _params=None, device=None, **kwargs):
        super().__init__(**kwargs)
        self.input_height, self.input_width, self.input_depth = (int(s) for s in in_shape)
        self.which_mask = which_mask
        self.num_res_blocks = num_res_blocks
        self.cardinality = cardinality
        self.kernel_size = kernel_size
        self.init = init
        self.LAYER_NORM = LAYER_NORM
        self.which_dilations = which_dilations
        self.device = torch.device(device) if device is not None else _default_device()
        self._owns_handle = _handle is None
        if _handle is None:
            h = c_void_p()
            dil = int_array(which_dilations)
            check(lib.cnf_coupling_create(int_array(in_shape), int(which_mask), int(num_res_blocks),
                                          int(cardinality), int(num_kernels), int(kernel_size),
                                          1 if LAYER_NORM else 0, dil, len(dil), byref(h)))
            _handle = h
        self._h = _handle
        info = _lib.coupling_info(self._h)
        self._info = info
        self.num_kernels = info.nk                           # halved for masks 0/1 (M:420-423)
        self.which_mask_complement = info.mask_complement    # M:426-433
        self.compressed_height, self.compressed_width, self.compressed_depth = info.h, info.w, info.c1
        self.uv2_depth = info.c2
        self._entries = _lib.coupling_entries(self._h, info.n_entries)
        self._net_stride = int(info.net_stride)
        if _params is None:
            _params = torch.zeros(int(info.param_count), dtype=torch.float32, device=self.device)
            self.params = _params
            self._initialise(initializers.get(init))
        else:
            self.params = _params
        self._ws = None
        self.model_A, self.model_b = _NetHandle(self, "A"), _NetHandle(self, "b")
        self.last_logdet_per_sample = None

    def __del__(self):
        try:
            if getattr(self, "_owns_handle", False) and self._h:
                lib.cnf_coupling_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def get_config(self):
        config = super().get_config()
        config.update({'input_height': self.input_height, 'input_width': self.input_width,
                       'input_depth': self.input_depth, 'which_mask': self.which_mask,
                       'num_res_blocks': self.num_res_blocks, 'cardinality': self.cardinality,
                       'kernel_size': self.kernel_size, 'init': self.init, 'LAYER_NORM': self.LAYER_NORM,
                       'which_dilations': self.which_dilations})
        return config

    def get_masked_compressed_shape(self):                   # M:474-498 (done by the planner)
        return self.compressed_height, self.compressed_width, self.compressed_depth

    # -- weights --------------------------------------------------------------------------------
    def _host_flat(self, fill):
        flat = np.zeros(2 * self._net_stride, np.float32)
        for n, net in enumerate(("A", "b")):
            base = n * self._net_stride
            for name, off, shape, role in self._entries:
                val = fill(net, name, shape, _ROLE[role])
                if val is None:
                    continue
                cnt = int(np.prod(shape))
                flat[base + off: base + off + cnt] = np.asarray(val, np.float32).reshape(-1)
        return flat

    def _initialise(self, init):
        def fill(net, name, shape, role):
            if role == "kernel":
                return init(shape)
            if role in ("gamma", "scale"):
                return np.ones(shape, np.float32)        # LN gamma 1; tanh scale 1 (M:109-112)
            return None                                  # biases / beta stay 0
        self.params.copy_(torch.from_numpy(self._host_flat(fill)))

    def weight_views(self, buf=None):
        """{'A': {name: tensor view}, 'b': {...}} — Keras-shaped views into the flat device buffer (or into
        `buf`, a tensor with the same layout, e.g. this layer's slice of the gradient buffer)."""
        buf = self.params if buf is None else buf
        out = {}
        for n, net in enumerate(("A", "b")):
            base = n * self._net_stride
            d = {}
            for name, off, shape, role in self._entries:
                if role == 4 and net == "b":
                    continue                             # net b has no tanh scale (M:1190-1204)
                cnt = int(np.prod(shape))
                v = buf[base + off: base + off + cnt]
                d[name] = v.view(()) if role == 4 else v.view(shape)
            out[net] = d
        return out

    def get_weights(self):
        return {net: {k: v.detach().cpu().numpy().copy() for k, v in d.items()}
                for net, d in self.weight_views().items()}

    def _random_state_spec(self):
        """per-element (std, is_uniform) of the "trained-like" random state (SURVEY 8d W-rand): kernels N(0, 1/fan_in), biases
        and beta N(0, 0.1), gamma and the tanh scale U[0.5, 1.5]; unused padding stays 0"""
        std = np.zeros(2 * self._net_stride, np.float32)
        uni = np.zeros(2 * self._net_stride, bool)
        for n, net in enumerate(("A", "b")):
            base = n * self._net_stride
            for name, off, shape, role in self._entries:
                if role == 4 and net == "b":
                    continue
                cnt = int(np.prod(shape))
                if role == 0:
                    std[base + off: base + off + cnt] = 1.0 / math.sqrt(int(np.prod(shape[:-1])))
                elif role in (2, 4):
                    uni[base + off: base + off + cnt] = True
                else:
                    std[base + off: base + off + cnt] = 0.1
        return std, uni

    def randomize_weights(self, generator=None):
        """"Trained-like" random state for benchmarks and stress tests (no reference counterpart), drawn on the device."""
        std, uni = self._random_state_spec()
        dev = self.params.device
        n = self.params.numel()
        p = torch.randn(n, device=dev, generator=generator) * torch.from_numpy(std).to(dev)
        u = torch.from_numpy(uni).to(dev)
        self.params.copy_(torch.where(u, 0.5 + torch.rand(n, device=dev, generator=generator), p))

    def resident_kernel_eligible(self):
        """True iff inference runs this layer as ONE activation-resident launch (csrc/fused_kernels.cu)."""
        return bool(lib.cnf_coupling_resident_eligible(self._h))

    def set_weights(self, weights):
        names = {net: set(d) for net, d in self.weight_views().items()}
        for net in ("A", "b"):
            missing = names[net] - set(weights[net])
            extra = set(weights[net]) - names[net]
            if missing or extra:
                raise ValueError(f"set_weights: net {net}: missing {sorted(missing)[:4]} / unexpected {sorted(extra)[:4]}")

        def fill(net, name, shape, role):
            if role == "scale" and net == "b":
                return np.ones(shape, np.float32)
            val = np.asarray(weights[net][name], np.float32)
            if val.size != int(np.prod(shape)):
                raise ValueError(f"set_weights: {net}.{name}: expected shape {shape}, got {val.shape}")
            return val
        self.params.copy_(torch.from_numpy(self._host_flat(fill)))

    # -- plumbing ---------------------------------------------------------------------------------
    def _workspace(self, B):
        need = int(lib.cnf_coupling_workspace_bytes(self._h, B))
        if self._ws is None or self._ws.numel() < need or self._ws.device != self.params.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.params.device)
        return self._ws

    def _check_uv(self, uv, name):
        uv = require_cuda(uv, name)
        want = (self.input_height, self.input_width, self.input_depth)
        if uv.dim() != 4 or tuple(uv.shape[1:]) != want:       # tf.ensure_shape (M:621, M:1276, M:1348)
            raise ValueError(f"{name}: shape {tuple(uv.shape)} is not compatible with [None, {want[0]}, {want[1]}, {want[2]}]")
        return uv

    def _nets(self, u1c):
        u1c = require_cuda(u1c, "u1_compressed")
        B = u1c.shape[0]
        A = torch.empty((B, self.compressed_height, self.compressed_width, self.uv2_depth),
                        dtype=torch.float32, device=u1c.device)
        b = torch.empty_like(A)
        br = Borrowed()
        check(lib.cnf_coupling_nets(self._h, br(u1c), br(self.params), br(A), br(b), br(self._workspace(B)),
                                    stream_ptr()))
        return A, b

    def coupling_function(self):
        """M:1076-1213: the two s/t networks of the layer as callables `(model_A, model_b)` on the compressed u1
        [B,h,w,c1] -> [B,h,w,c2].  The reference builds and returns two Keras functional models here; the networks of
        this layer already exist (flat parameter buffer + libcnf kernels), so the handles are returned, with
        `.get_weights()` like the Keras models.  model_A includes the `w * tanh(.)` scaling layer (M:1198)."""
        return self.model_A, self.model_b

    def residual_block(self, x, block, net="both"):
        """Residual block `block` (F:501-627) of this layer's networks on stand-alone activations.
        x: [2,B,h,w,nk] (net A, net b) with net="both", or [B,h,w,nk] for net "A" / "b"."""
        if net == "both":
            x2 = require_cuda(x, "x")
        else:
            x1 = require_cuda(x, "x")
            x2 = torch.stack([x1, x1]).contiguous()
        if x2.dim() != 5:
            raise ValueError(f"x: expected [2,B,h,w,nk] (or [B,h,w,nk] for one net), got {tuple(x.shape)}")
        out = torch.empty_like(x2)
        br = Borrowed()
        check(lib.cnf_residual_block(self._h, int(block), br(x2), br(self.params), br(out),
                                     br(self._workspace(x2.shape[1])), stream_ptr()))
        return out if net == "both" else out[0 if net == "A" else 1]

    def A_wrapper(self, A_input):                            # M:452-461
        return self.model_A(A_input)

    def b_wrapper(self, b_input):                            # M:463-472
        return self.model_b(b_input)

    # -- masks (M:500-761, M:763-1073) ------------------------------------------------------------
    def mask(self, uv, which_mask_index, compress):
        uv = self._check_uv(uv, "uv")
        B, H, W, D = uv.shape
        if compress:
            if which_mask_index in (0, 1):
                shape = (B, H // 2, W // 2, 2 * D)
            else:
                shape = (B, H, W, (D + 1) // 2 if which_mask_index == 2 else D // 2)
        else:
            shape = (B, H, W, D)
        out = torch.empty(shape, dtype=torch.float32, device=uv.device)
        br = Borrowed()
        check(lib.cnf_mask(br(uv), int(which_mask_index), 1 if compress else 0, br(out), stream_ptr()))
        return out

    def decompress_mask(self, uv_masked_compressed, which_mask_index, uv_shape_OUTPUT):
        uvc = require_cuda(uv_masked_compressed, "uv_masked_compressed")
        if which_mask_index in (0, 1):
            assert uvc.shape[3] % 2 == 0, \
                'The compressed, checkerboard-masked u/v should always have an even number of channels.'
        shape = (uvc.shape[0],) + tuple(int(s) for s in uv_shape_OUTPUT[1:])
        out = torch.empty(shape, dtype=torch.float32, device=uvc.device)
        br = Borrowed()
        check(lib.cnf_decompress_mask(br(uvc), int(which_mask_index), br(out), stream_ptr()))
        return out

    # -- coupling laws on compressed halves (M:1215-1253); elementwise torch, API parity only ------
    def forward_coupling_law(self, exp_A_u1, b_u1, u2_compressed):
        return exp_A_u1 * u2_compressed + b_u1

    def inverse_coupling_law(self, inv_exp_A_v1, b_v1, v2_compressed):
        return inv_exp_A_v1 * (v2_compressed - b_v1)

    def set_fusion(self, enable):
        """Inference path selection (no reference counterpart): one activation-resident launch per layer when the
        layer fits it (default) or the layer-per-kernel path."""
        check(lib.cnf_coupling_set_fusion(self._h, 1 if enable else 0))

    def set_kernel_paths(self, excluded):
        """`excluded`: CNF_PATH_* bits (see _lib / include/cnf.h) naming the kernel families this layer must not use;
        0 = the fastest eligible kernel for every stage.  All choices give the same results to fp32 rounding."""
        check(lib.cnf_coupling_set_kernel_paths(self._h, int(excluded)))

    # -- the layer (M:1258-1328, M:1333-1394) -------------------------------------------------------
    def forward_and_Jacobian(self, u, sum_log_detJ, zy):
        u = self._check_uv(u, "u")
        B = u.shape[0]
        v = torch.empty_like(u)
        ld = torch.empty(B, dtype=torch.float32, device=u.device)
        br = Borrowed()
        check(lib.cnf_coupling_forward(self._h, br(u), br(self.params), br(v), br(ld), br(self._workspace(B)),
                                       stream_ptr()))
        self.last_logdet_per_sample = ld
        sum_log_detJ = sum_log_detJ + ld.mean()              # M:1325-1326 (batch mean inside the layer, Q1)
        return v, sum_log_detJ, zy

    def backward(self, v, zy):
        v = self._check_uv(v, "v")
        B = v.shape[0]
        u = torch.empty_like(v)
        br = Borrowed()
        check(lib.cnf_coupling_backward(self._h, br(v), br(self.params), br(u), br(self._workspace(B)),
                                        stream_ptr()))
        return u, zy


###############################################################################

class cFlow:
    """Conditional RealNVP flow (M:1396-1904): same constructor, attributes and calls as the reference."""

    def __init__(self, io_shape, x_d, squeeze_factor_block_list, ResNeXt_block_list, num_kernels_list,
                 cardinality_list, lambda_y=100, ksize=3, LAYER_NORM=True, DILATIONS=True,
                 init=None, device=None):
        self.io_shape = io_shape
        self.x_d = x_d
        self.squeeze_factor_block_list = squeeze_factor_block_list
        self.ResNeXt_block_list = ResNeXt_block_list
        self.num_kernels_list = num_kernels_list
        self.cardinality_list = cardinality_list
        self.lambda_y = lambda_y
        self.ksize = ksize
        self.LAYER_NORM = LAYER_NORM
        self.DILATIONS = DILATIONS
        self.init = init if init is not None else initializers.Orthogonal(gain=0.1)   # M:1442
        self.device = torch.device(device) if device is not None else _default_device()
        self.optimizer = None
        self._plan = None

        assert len(self.squeeze_factor_block_list) == len(self.ResNeXt_block_list) == \
            len(self.num_kernels_list) == len(self.cardinality_list), \
            'squeeze_factor_block_list, ResNeXt_block_list, num_kernels_list, and cardinality_list must all have the same length.'
        self.num_coupling_blocks = len(self.squeeze_factor_block_list)

        h = c_void_p()
        rc = lib.cnf_plan_create(int_array(io_shape), int(x_d), self.num_coupling_blocks,
                                 int_array(squeeze_factor_block_list), int_array(ResNeXt_block_list),
                                 int_array(num_kernels_list), int_array(cardinality_list), float(lambda_y),
                                 int(ksize), 1 if LAYER_NORM else 0, 1 if DILATIONS else 0, byref(h))
        if rc == _lib.CNF_ERR_UNSUPPORTED and not DILATIONS:
            # the reference dereferences self.dilations_list, which only exists when DILATIONS (M:1553)
            raise AttributeError(lib.cnf_last_error().decode())
        check(rc)
        self._plan = h
        info = _lib.PlanInfo()
        check(lib.cnf_plan_get_info(self._plan, byref(info)))
        self._info = info

        # derived attributes, same names/types as the reference (M:1493-1617)
        blocks = []
        for i in range(info.n_blocks):
            b = _lib.BlockInfo()
            check(lib.cnf_plan_block_info(self._plan, i, byref(b)))
            blocks.append(b)
        self.scale_list = np.array([b.scale for b in blocks])
        self.num_prev_factors_list = np.array([b.num_prev_factors for b in blocks])
        self.io_shape_list = np.array([[b.H, b.W, b.D] for b in blocks])
        self.u1_mask_indices = [[0, 1, 2, 3] for _ in blocks]
        self.dilations_list = [{'checkerboard': [b.checkerboard[j] for j in range(b.n_checkerboard)],
                                'channelwise': [b.channelwise[j] for j in range(b.n_channelwise)]}
                               for b in blocks]
        self.distribution = _StandardNormalDiag(self.x_d, self.device)               # M:1621-1623

        # one flat parameter buffer; every coupling layer views its slice
        self.params = torch.zeros(int(info.param_count), dtype=torch.float32, device=self.device)
        init_fn = initializers.get(self.init)
        self.layers_list = []
        self.squeeze_factor_layers_list = []
        self.coupling_layers = []
        host = np.zeros(int(info.param_count), np.float32)
        for idx in range(info.n_layers):
            kind, aux = c_int(), c_int()
            check(lib.cnf_plan_layer(self._plan, idx, byref(kind), byref(aux)))
            if kind.value == 0:
                li = aux.value
                ch = c_void_p(lib.cnf_plan_coupling(self._plan, li))
                ci = _lib.coupling_info(ch)
                off = int(lib.cnf_plan_coupling_param_offset(self._plan, li))
                blk = li // 4
                layer = coupling_layer(in_shape=[ci.H, ci.W, ci.D], which_mask=ci.mask,
                                       num_res_blocks=self.ResNeXt_block_list[blk],
                                       cardinality=self.cardinality_list[blk],
                                       num_kernels=self.num_kernels_list[blk], kernel_size=self.ksize,
                                       init=self.init, LAYER_NORM=self.LAYER_NORM,
                                       which_dilations=[ci.dilation[j] for j in range(ci.n_branches)],
                                       _handle=ch, _params=self.params[off: off + int(ci.param_count)],
                                       device=self.device)

                def fill(net, name, shape, role, _init=init_fn):
                    if role == "kernel":
                        return _init(shape)
                    if role in ("gamma", "scale"):
                        return np.ones(shape, np.float32)
                    return None
                host[off: off + int(ci.param_count)] = layer._host_flat(fill)
                self.layers_list.append(layer)
                self.coupling_layers.append(layer)
            elif kind.value == 1:
                layer = squeeze_layer()
                self.layers_list.append(layer)
                self.squeeze_factor_layers_list.append(layer)
            else:
                layer = factor_out_zy_layer(aux.value)
                self.layers_list.append(layer)
                self.squeeze_factor_layers_list.append(layer)
        self.params.copy_(torch.from_numpy(host))

        self.loss_tracker = _Mean('loss')
        self.z_loss_tracker = _Mean('z_loss')
        self.y_loss_tracker = _Mean('y_loss')
        self.detJ_loss_tracker = _Mean('detJ_loss')
        self._ws = None
        self.last_logdet_per_sample = None
        self.last_per_sample = None

    def __del__(self):
        try:
            if self._plan:
                lib.cnf_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    # -- keras-ish surface --------------------------------------------------------------------------
    @property
    def metrics(self):
        return [self.loss_tracker, self.z_loss_tracker, self.y_loss_tracker, self.detJ_loss_tracker]

    @property
    def trainable_variables(self):
        return [self.params]

    def compile(self, optimizer=None):
        self.optimizer = optimizer

    def randomize_weights(self, seed=0):
        """"Trained-like" random weights for every coupling layer (see coupling_layer.randomize_weights)."""
        g = torch.Generator(device=self.params.device)
        g.manual_seed(int(seed))
        for layer in self.coupling_layers:
            layer.randomize_weights(g)

    def launches_per_pass(self):
        """kernel launches of one pass of the flow (forward or inverse) through the coupling layers"""
        return sum(1 if l.resident_kernel_eligible() else 2 + 3 * l._info.R for l in self.coupling_layers)

    def set_fusion(self, enable):
        """Inference path selection for every coupling layer (see coupling_layer.set_fusion)."""
        check(lib.cnf_plan_set_fusion(self._plan, 1 if enable else 0))

    def set_kernel_paths(self, excluded):
        """Kernel-family exclusion bits for every coupling layer (see coupling_layer.set_kernel_paths)."""
        check(lib.cnf_plan_set_kernel_paths(self._plan, int(excluded)))

    def count_params(self):
        n = 0
        for layer in self.coupling_layers:
            for net, d in layer.weight_views().items():
                n += sum(v.numel() for v in d.values())
        return n

    def get_weights(self):
        return [layer.get_weights() for layer in self.coupling_layers]

    def set_weights(self, weights):
        if len(weights) != len(self.coupling_layers):
            raise ValueError(f"set_weights: expected {len(self.coupling_layers)} coupling layers, got {len(weights)}")
        for layer, w in zip(self.coupling_layers, weights):
            layer.set_weights(w)

    def save_weights(self, path):
        flat = {}
        for i, w in enumerate(self.get_weights()):
            for net in ("A", "b"):
                for k, v in w[net].items():
                    flat[f"layer{i}.{net}.{k}"] = v
        np.savez(path, **flat)

    def load_weights(self, path):
        data = np.load(path)
        weights = [{"A": {}, "b": {}} for _ in self.coupling_layers]
        for key in data.files:
            layer, net, name = key.split(".", 2)
            weights[int(layer[5:])][net][name] = data[key]
        self.set_weights(weights)

    def _workspace(self, B):
        """s/t-net workspace of the CURRENT stream: calls issued on different streams (log_loss_and_sample) may run
        concurrently on the device, so each stream owns its buffer."""
        need = int(lib.cnf_plan_workspace_bytes(self._plan, B))
        if self._ws is None:
            self._ws = {}
        key = torch.cuda.current_stream().cuda_stream if self.params.is_cuda else 0
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need or ws.device != self.params.device:
            self._ws[key] = ws = torch.empty(need, dtype=torch.uint8, device=self.params.device)
        return ws

    def _check_io(self, t, name):
        t = require_cuda(t, name)
        want = tuple(int(s) for s in self.io_shape)
        if t.dim() != 4 or tuple(t.shape[1:]) != want:
            raise ValueError(f"{name}: shape {tuple(t.shape)} is not compatible with [None, {want[0]}, {want[1]}, {want[2]}]")
        return t

    # -- cFlow.call (M:1723-1798) ---------------------------------------------------------------------
    def call(self, uv, direction=-1):
        if direction == 1:
            xy = self._check_io(uv, "xy")
            B = xy.shape[0]
            zy = torch.empty_like(xy)
            ld = torch.empty(B + 1, dtype=torch.float32, device=xy.device)   # [B] per sample + their batch mean
            br = Borrowed()
            check(lib.cnf_flow_forward(self._plan, br(xy), br(self.params), br(zy), br(ld),
                                       br(self._workspace(B)), stream_ptr()))
            self.last_logdet_per_sample = ld[:B]
            return zy, ld[B]                                 # Q1: scalar = batch mean of per-sample log-dets (same launch)
        elif direction == -1:
            zy = self._check_io(uv, "zy")
            B = zy.shape[0]
            xy = torch.empty_like(zy)
            br = Borrowed()
            check(lib.cnf_flow_inverse(self._plan, br(zy), br(self.params), br(xy), br(self._workspace(B)),
                                       stream_ptr()))
            return xy
        # the reference falls through both branches and returns None for any other direction (M:1743, M:1774)
        return None

    __call__ = call

    def log_loss_and_sample(self, xy, zy):
        """log_loss(xy) and call(zy, -1) -- the two halves of a likelihood-evaluation + sampling step, which do not depend
        on each other -- issued on TWO CUDA streams.  Every s/t-net kernel is a persistent grid of one CTA per SM whose
        CTAs finish at different times (20 to 22 tiles each) and start with a prologue (resident weights, gamma / beta, tensor
        memory); with a second, independent kernel queue the SMs that one pass leaves idle at every kernel boundary run
        the other pass's CTAs.  Same kernels, same results as the two calls in sequence (each stream has its own workspace).
        Returns ((loss, z_loss, y_loss, detJ_loss), samples); both are ready on the current stream."""
        main = torch.cuda.current_stream()
        side = getattr(self, '_side_stream', None)
        if side is None or side.device != self.params.device:
            self._side_stream = side = torch.cuda.Stream(device=self.params.device)
        side.wait_stream(main)                       # zy (and the weights) were produced on the current stream
        with torch.cuda.stream(side):
            samples = self.call(zy, -1)
        four = self.log_loss(xy)
        main.wait_stream(side)
        samples.record_stream(main)                  # allocated under the side stream, used by the caller on this one
        return four, samples

    # -- cFlow.log_loss (M:1800-1848) -------------------------------------------------------------------
    def log_loss(self, xy):
        xy = self._check_io(xy, "xy")
        B = xy.shape[0]
        dev = xy.device
        zy = torch.empty_like(xy)
        Bp = (B + 3) & ~3                                     # keep every row 16-byte aligned
        pers = torch.empty((3, Bp), dtype=torch.float32, device=dev)
        loss4 = torch.empty(4, dtype=torch.float32, device=dev)
        ll_z, ll_y, ld = pers[0, :B], pers[1, :B], pers[2, :B]
        br = Borrowed()
        check(lib.cnf_flow_log_loss(self._plan, br(xy), br(self.params), br(zy), br(ll_z), br(ll_y), br(ld),
                                    br(loss4), br(self._workspace(B)), stream_ptr()))
        self.last_logdet_per_sample = ld
        self.last_per_sample = {'ll_z': ll_z, 'll_y': ll_y, 'logdet': ld, 'zy': zy}
        return loss4[0], loss4[1], loss4[2], loss4[3]

    def bits_per_dim(self, z_loss, detJ_loss):
        """derived metric; the reference never computes it."""
        H, W, _ = (int(s) for s in self.io_shape)
        return (float(z_loss) + float(detJ_loss)) / (H * W * self.x_d * math.log(2.0))

    # -- train / test steps (M:1850-1904) -----------------------------------------------------------------
    def _update_trackers(self, four):
        vals = [float(v) for v in torch.stack(list(four)).cpu()]
        for tr, v in zip(self.metrics, vals):
            tr.update_state(v)
        return {'loss': self.loss_tracker.result(), 'z_loss': self.z_loss_tracker.result(),
                'y_loss': self.y_loss_tracker.result(), 'detJ_loss': self.detJ_loss_tracker.result()}

    def test_step(self, xy):
        return self._update_trackers(self.log_loss(xy))

    # -- gradients of log_loss (the tf.GradientTape block of train_step, M:1863-1871) --------------------------
    #: True: keep only the per-layer flow states in the forward pass and re-compute each layer's s/t-net activations during
    #: the backward pass (same gradients, ~1/n_layers of the activation memory, one extra forward pass)
    recompute_activations = False
    #: True: keep NOTHING in the forward pass; the backward pass recovers each layer's input state from its output with the
    #: inverse law (M:1333-1394) and re-computes the activations from it (least memory; gradients carry the flow's fp32
    #: round-trip error, see cnf.h)
    recover_states_by_inverse = False

    def _train_fns(self):
        if self.recover_states_by_inverse:
            return lib.cnf_plan_train_workspace_bytes_invert, lib.cnf_flow_loss_and_grad_invert
        if self.recompute_activations:
            return lib.cnf_plan_train_workspace_bytes_recompute, lib.cnf_flow_loss_and_grad_recompute
        return lib.cnf_plan_train_workspace_bytes, lib.cnf_flow_loss_and_grad

    def _train_workspace(self, B):
        need = int(self._train_fns()[0](self._plan, B))
        ws = getattr(self, '_train_ws', None)
        if ws is None or ws.numel() < need or ws.device != self.params.device:
            self._train_ws = ws = torch.empty(need, dtype=torch.uint8, device=self.params.device)
        return ws

    def _grad_buffer(self):
        if getattr(self, '_grads', None) is None or self._grads.device != self.params.device:
            self._grads = torch.empty_like(self.params)
        return self._grads

    def loss_and_grad(self, xy, on_layer_grads=None):
        """log_loss(xy) plus dloss/dparams as ONE flat tensor with the layout of self.params (hand-written
        backward kernels).  Returns ((loss, z_loss, y_loss, detJ_loss), grads).
        on_layer_grads(layer, offset, count), if given, is called on the host in backward order as soon as the kernels
        that complete grads[offset: offset + count] of coupling layer `layer` are enqueued on the current stream
        (cnf_flow_loss_and_grad_hooked): the hook of the overlapped gradient all-reduce, sharding.BucketedGradAllReduce."""
        xy = self._check_io(xy, "xy")
        B = xy.shape[0]
        dev = xy.device
        zy = torch.empty_like(xy)
        Bp = (B + 3) & ~3
        pers = torch.empty((3, Bp), dtype=torch.float32, device=dev)
        loss4 = torch.empty(4, dtype=torch.float32, device=dev)
        ll_z, ll_y, ld = pers[0, :B], pers[1, :B], pers[2, :B]
        grads = self._grad_buffer()
        br = Borrowed()
        if on_layer_grads is None:
            fn = self._train_fns()[1]
            check(fn(self._plan, br(xy), br(self.params), br(grads), br(zy), br(ll_z), br(ll_y), br(ld), br(loss4),
                     br(self._train_workspace(B)), stream_ptr()))
        else:
            mode = 2 if self.recover_states_by_inverse else 1 if self.recompute_activations else 0
            cb = LAYER_GRADS_READY_FN(lambda _user, layer, off, count: on_layer_grads(layer, off, count))
            check(lib.cnf_flow_loss_and_grad_hooked(self._plan, br(xy), br(self.params), br(grads), br(zy), br(ll_z), br(ll_y),
                                                    br(ld), br(loss4), br(self._train_workspace(B)), stream_ptr(), mode,
                                                    cb, None))
        self.last_logdet_per_sample = ld
        self.last_per_sample = {'ll_z': ll_z, 'll_y': ll_y, 'logdet': ld, 'zy': zy}
        return (loss4[0], loss4[1], loss4[2], loss4[3]), grads

    def grad_views(self, grads=None):
        """the flat gradient buffer as the same named views get_weights() uses: [{'A': {...}, 'b': {...}}, ...]"""
        grads = self._grads if grads is None else grads
        out = []
        for layer in self.coupling_layers:
            off = layer.params.storage_offset() - self.params.storage_offset()
            out.append(layer.weight_views(grads[off: off + layer.params.numel()]))
        return out

    # data-parallel train_step: gradient buckets of at least this many bytes are all-reduced while the backward pass is
    # still running (sharding.BucketedGradAllReduce); 0 / None: one flat all-reduce after the backward pass
    grad_bucket_bytes = 8 << 20

    def train_step(self, xy):
        """cFlow.train_step (M:1850-1880): gradients of log_loss, optimizer.apply_gradients, metric trackers.
        Data-parallel (torch.distributed initialised): the flat gradient buffer is sum all-reduced over NCCL in buckets of
        coupling layers, each issued as soon as the backward pass has completed it (sharding.BucketedGradAllReduce,
        SURVEY 8e; grad_bucket_bytes = 0: one all-reduce after the backward pass, sharding.allreduce_mean_gradients);
        the reported metrics are this rank's shard's."""
        if self.optimizer is None:
            raise RuntimeError("train_step: call model.compile(optimizer=Adam(...)) first")   # keras raises too
        from .sharding import BucketedGradAllReduce, allreduce_mean_gradients, sync_replicas
        if not getattr(self, '_replicas_synced', False):
            # data-parallel replicas must start from the same weights / optimizer state (rank 0's); single process: no-op
            sync_replicas(self)
            self._replicas_synced = True
        parallel = torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size() > 1
        if parallel and self.grad_bucket_bytes:
            red = BucketedGradAllReduce(self._grad_buffer(), xy.shape[0], self.grad_bucket_bytes)
            four, grads = self.loss_and_grad(xy, on_layer_grads=red.layer_ready)
            red.finish()
        else:
            four, grads = self.loss_and_grad(xy)
            allreduce_mean_gradients(grads, xy.shape[0])
        self.optimizer.apply_gradients(self.params, grads)
        return self._update_trackers(four)

    def fit(self, batches, epochs=1, verbose=0):
        """thin replacement of the keras fit loop (C:617-636, P:136-145): `batches` is an iterable of xy tensors
        (re-iterated every epoch).  Returns {'loss': [...], 'z_loss': [...], ...} per epoch, like History.history."""
        history = {m.name: [] for m in self.metrics}
        for ep in range(epochs):
            for m in self.metrics:
                m.reset_state()
            logs = None
            for xy in batches:
                logs = self.train_step(xy)
            if logs is None:
                raise ValueError("fit: empty dataset")
            for k, v in logs.items():
                history[k].append(v)
            if verbose:
                print(f"epoch {ep + 1}/{epochs} " + " ".join(f"{k}={v:.5f}" for k, v in logs.items()))
        return history


class Adam:
    """tf.keras.optimizers.Adam(learning_rate) as the reference compiles it (C:567, P:130): defaults lr 1e-3,
    beta_1 0.9, beta_2 0.999, epsilon 1e-7, no amsgrad.  One fused kernel over the flat parameter buffer."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate = learning_rate
        self.beta_1 = beta_1
        self.beta_2 = beta_2
        self.epsilon = epsilon
        self.iterations = 0
        self._m = None
        self._v = None

    def apply_gradients(self, params, grads, grad_scale=1.0):
        if self._m is None or self._m.shape != params.shape or self._m.device != params.device:
            self._m = torch.zeros_like(params)
            self._v = torch.zeros_like(params)
        self.iterations += 1
        br = Borrowed()
        check(lib.cnf_adam_step(br(params), br(grads), br(self._m), br(self._v), self.iterations,
                                float(self.learning_rate), float(self.beta_1), float(self.beta_2),
                                float(self.epsilon), float(grad_scale), stream_ptr()))
