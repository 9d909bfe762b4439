"""Oracle: torch-CPU functional restatement of the conv cINN hot path (fp32 or fp64).

Follows /root/reference:
  * s/t networks  conv_cINN_make_model.py M:1076-1213 (coupling_function),
                  conv_cINN_base_functions.py F:330-362 (add_common_layers),
                  F:364-413 (grouped_convolution), F:501-627 (dilated_residual_block),
                  M:97-122 (tanh_scaling_layer)
  * coupling      M:1215-1253 (laws), M:1258-1328 (forward_and_Jacobian), M:1333-1394 (backward)
  * flow          M:1723-1798 (cFlow.call), M:1800-1848 (log_loss)
Keras semantics encoded explicitly: NHWC, HWIO kernels, cross-correlation, stride 1,
padding='same' (= dilation*(k-1)/2 each side for odd k), LeakyReLU slope 0.3,
LayerNormalization over the flattened h*w*C vector with per-element gamma/beta,
biased variance, eps 1e-3; MultivariateNormalDiag(0, I).log_prob.
Index permutations go through the literal transcription in `masks_np`.
TEST INFRASTRUCTURE (see oracle/__init__.py); also the `cpu_baseline` port timed by bench.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import masks_np
from .planner import plan_flow

LRELU = 0.3      # keras LeakyReLU() default alpha (F:345)
LN_EPS = 1e-3    # keras LayerNormalization default epsilon (F:358)


def _t(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype)


def conv2d_same(x, kernel, bias, dilation=1):
    """Keras Conv2D(padding='same', strides 1) on NHWC torch tensor; kernel HWIO."""
    k = kernel.shape[0]
    assert k % 2 == 1, "oracle restates odd kernel sizes only"
    pad = dilation * (k - 1) // 2
    w = kernel.permute(3, 2, 0, 1).contiguous()              # OIHW
    y = F.conv2d(x.permute(0, 3, 1, 2).contiguous(), w, bias, stride=1, padding=pad, dilation=dilation)
    return y.permute(0, 2, 3, 1)


def common_layers(y, gamma, beta, ln=True):
    """F:330-362: LeakyReLU then (optionally) LayerNorm over the flattened sample."""
    y = F.leaky_relu(y, LRELU)
    if ln:
        B, h, w, c = y.shape
        f = y.reshape(B, h * w * c)
        f = F.layer_norm(f, (h * w * c,), gamma, beta, LN_EPS)
        y = f.reshape(B, h, w, c)
    return y


#: How the `Lambda(lambda z: z[:, :, :, j * _d:j * _d + _d])` slices of F:402 bind `j` (oracle/tf_shim/README.md).
#: False: every group reads its own slice (the lambda runs once, while the graph is built: TF 1.x graph-mode Keras; SURVEY
#: quirk Q11; what the CUDA kernels implement).  True: tf.keras >= 2.4 re-runs the lambda on every model call, when the loop
#: variable holds its final value, so EVERY group reads the last group's channels.  Both are pinned against the reference's
#: own source (tests/golden/refsrc_*.npz); which one a real TensorFlow 2.7 produces must be decided on a TensorFlow box
#: (tools/tf_dump_reference.py writes the probe).
LAMBDA_LATE_BINDING = False


def grouped_convolution(y, P, prefix, nb_channels, ksize, dilation, cardinality):
    """F:364-413, literal: one Conv2D per group over channel slice [j*_d, (j+1)*_d)."""
    if cardinality == 1:
        return conv2d_same(y, P[f'{prefix}.g0.kernel'], P[f'{prefix}.g0.bias'], dilation)
    assert not nb_channels % cardinality
    _d = nb_channels // cardinality
    groups = []
    for j in range(cardinality):
        js = cardinality - 1 if LAMBDA_LATE_BINDING else j
        grp = y[..., js * _d:js * _d + _d]
        groups.append(conv2d_same(grp, P[f'{prefix}.g{j}.kernel'], P[f'{prefix}.g{j}.bias'], dilation))
    return torch.cat(groups, dim=-1)


def dilated_residual_block(y, P, r, L):
    """F:501-627 with nb_channels_in = nb_channels_out = nk, strides 1, no projection."""
    ln = L['ln']
    shortcut = y
    y = common_layers(y, P.get(f'rb{r}.ln1.gamma'), P.get(f'rb{r}.ln1.beta'), ln)
    y = conv2d_same(y, P[f'rb{r}.pw1.kernel'], P[f'rb{r}.pw1.bias'])
    y = common_layers(y, P.get(f'rb{r}.ln2.gamma'), P.get(f'rb{r}.ln2.beta'), ln)
    outs = []
    for br in L['branches']:
        outs.append(grouped_convolution(y, P, f"rb{r}.gc.d{br['dilation']}", br['channels'],
                                        L['ksize'], br['dilation'], L['cardinality']))
    y = torch.cat(outs, dim=-1) if len(outs) > 1 else outs[0]
    y = common_layers(y, P.get(f'rb{r}.ln3.gamma'), P.get(f'rb{r}.ln3.beta'), ln)
    y = conv2d_same(y, P[f'rb{r}.pw2.kernel'], P[f'rb{r}.pw2.bias'])
    return shortcut + y


def st_net(u1c, P, L, is_A):
    """M:1106-1205: stem conv -> R residual blocks -> LReLU/LN -> head conv (-> w*tanh for A)."""
    x = conv2d_same(u1c, P['stem.kernel'], P['stem.bias'])
    for r in range(L['R']):
        x = dilated_residual_block(x, P, r, L)
    x = common_layers(x, P.get('lnf.gamma'), P.get('lnf.beta'), L['ln'])
    x = conv2d_same(x, P['head.kernel'], P['head.bias'])
    if is_A:
        x = P['tanh_scale'] * torch.tanh(x)                  # M:1198, M:114-116
    return x


class FlowOracle:
    """Restated cFlow.  `weights` is a list (one entry per coupling layer, in layers_list
    order) of {'A': {name: ndarray}, 'b': {name: ndarray}}."""

    def __init__(self, io_shape, x_d, squeeze_factor_block_list, ResNeXt_block_list,
                 num_kernels_list, cardinality_list, lambda_y=100, ksize=3, LAYER_NORM=True,
                 DILATIONS=True, weights=None, dtype=torch.float32):
        self.plan = plan_flow(io_shape, x_d, squeeze_factor_block_list, ResNeXt_block_list,
                              num_kernels_list, cardinality_list, ksize, DILATIONS)
        self.x_d, self.lambda_y, self.dtype = x_d, lambda_y, dtype
        self.layers = self.plan['layers']
        for L in self.layers:
            if L['type'] == 'coupling':
                L['ln'] = LAYER_NORM
        self.coupling = [L for L in self.layers if L['type'] == 'coupling']
        self.sf_layers = [L for L in self.layers if L['type'] != 'coupling']
        if weights is not None:
            self.set_weights(weights)

    def set_weights(self, weights):
        assert len(weights) == len(self.coupling)
        self.W = [{net: {k: _t(v, self.dtype) for k, v in w[net].items()} for net in ('A', 'b')}
                  for w in weights]

    # -- one coupling layer ------------------------------------------------------------
    def _nets(self, li, u1c):
        L = self.coupling[li]
        x = _t(u1c, self.dtype)
        with torch.no_grad():
            A = st_net(x, self.W[li]['A'], L, True)
            b = st_net(x, self.W[li]['b'], L, False)
        return A, b

    def coupling_forward(self, li, u):
        """M:1258-1328.  Returns v (ndarray) and the PER-SAMPLE log-det vector (torch)."""
        L = self.coupling[li]
        m, mc = L['mask'], L['mask_complement']
        v1 = masks_np.mask(u, m, False)
        u1c = masks_np.mask(u, m, True)
        u2c = _t(masks_np.mask(u, mc, True), self.dtype)
        A, b = self._nets(li, u1c)
        v2c = torch.exp(A) * u2c + b                         # M:1307-1312, M:1230-1231
        v2 = masks_np.decompress_mask(v2c.numpy(), mc, u.shape)
        v = v1 + v2                                          # M:1320
        return v, A.sum(dim=(1, 2, 3))                       # M:1323

    def coupling_backward(self, li, v):
        """M:1333-1394."""
        L = self.coupling[li]
        m, mc = L['mask'], L['mask_complement']
        u1 = masks_np.mask(v, m, False)
        v1c = masks_np.mask(v, m, True)
        v2c = _t(masks_np.mask(v, mc, True), self.dtype)
        A, b = self._nets(li, v1c)
        u2c = torch.reciprocal(torch.exp(A)) * (v2c - b)     # M:1379, M:1250-1251
        u2 = masks_np.decompress_mask(u2c.numpy(), mc, v.shape)
        return u1 + u2

    # -- the flow ----------------------------------------------------------------------
    def call(self, uv, direction=-1):
        """M:1723-1798.  direction=+1 returns (zy, log_detJ scalar, per-sample log-det);
        the third element is a superset of the reference's return (Q1)."""
        np_dtype = np.float64 if self.dtype == torch.float64 else np.float32
        uv = np.asarray(uv, dtype=np_dtype)
        if direction == 1:
            log_detJ = torch.zeros((), dtype=self.dtype)
            per_sample = torch.zeros(uv.shape[0], dtype=self.dtype)
            zy = None
            ci = 0
            for L in self.layers:
                if L['type'] == 'coupling':
                    uv, ld = self.coupling_forward(ci, uv)
                    log_detJ = log_detJ + ld.mean()          # M:1325-1326 (batch mean per layer)
                    per_sample = per_sample + ld
                    ci += 1
                elif L['type'] == 'squeeze':
                    uv, zy = masks_np.squeeze_forward(uv, zy)
                else:
                    uv, zy = masks_np.factor_forward(uv, zy)
            if len(self.sf_layers) == 0:                      # M:1757-1770
                vu = uv
            else:
                zy = np.concatenate([zy, uv], axis=3)
                vu = None
                for L in reversed(self.sf_layers):
                    if L['type'] == 'squeeze':
                        vu, zy = masks_np.squeeze_backward(vu, zy)
                    else:
                        vu, zy = masks_np.factor_backward(vu, zy, L['num_prev_factors'])
            return vu, log_detJ, per_sample
        assert direction == -1                                # M:1774-1798
        zy = None
        for L in self.sf_layers:
            if L['type'] == 'squeeze':
                uv, zy = masks_np.squeeze_forward(uv, zy)
            else:
                uv, zy = masks_np.factor_forward(uv, zy)
        vu = uv
        ci = len(self.coupling) - 1
        for L in reversed(self.layers):
            if L['type'] == 'coupling':
                vu = self.coupling_backward(ci, vu)
                ci -= 1
            elif L['type'] == 'squeeze':
                vu, zy = masks_np.squeeze_backward(vu, zy)
            else:
                vu, zy = masks_np.factor_backward(vu, zy, L['num_prev_factors'])
        return vu

    def log_loss(self, xy):
        """M:1800-1848.  Returns the reference 4-tuple plus per-sample (ll_z, ll_y, logdet)."""
        x_d = self.x_d
        np_dtype = np.float64 if self.dtype == torch.float64 else np.float32
        xy = np.asarray(xy, dtype=np_dtype)
        y_prime = _t(xy[..., x_d:], self.dtype)
        zy, log_detJ, ld_ps = self.call(xy, 1)
        zy = _t(zy, self.dtype)
        z, y = zy[..., :x_d], zy[..., x_d:]
        log_prob = -0.5 * (z * z).sum(-1) - 0.5 * x_d * math.log(2.0 * math.pi)   # M:1621-1623
        ll_z = log_prob.sum(dim=(1, 2))                       # M:1832
        ll_y = -self.lambda_y * (y - y_prime).abs().sum(dim=(1, 2, 3))              # M:1836-1838
        ll = (ll_z + ll_y).mean() + log_detJ                  # M:1840-1842
        four = (-ll, -ll_z.mean(), -ll_y.mean(), -log_detJ)   # M:1848
        return tuple(float(t) for t in four), {'ll_z': ll_z.numpy(), 'll_y': ll_y.numpy(),
                                               'logdet': ld_ps.numpy(), 'zy': zy.numpy()}


def bits_per_dim(z_loss, detJ_loss, H, W, x_d):
    """the reference never computes bits/dim; derived definition."""
    return (z_loss + detJ_loss) / (H * W * x_d * math.log(2.0))
