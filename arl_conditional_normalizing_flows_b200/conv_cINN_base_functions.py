"""Mirror of the hot-path part of the reference's `conv_cINN_base_functions.py` (F).

Only `dilated_residual_block` (F:501-627) with its helpers `grouped_convolution` (F:364-413) and
`add_common_layers` (F:330-362) is on the path SURVEY §8 names; `conv_cINN_make_model` imports it
(M:24).  In the reference it is a Keras functional-graph builder that is traced once while a
`coupling_layer` builds its two s/t models; here that construction is done by the C planner
(`cnf_coupling_create`: channel split per dilation F:577-588, groups F:396-411) and the block runs
inside the fused kernels of csrc/stnet_kernels.cu.

The data helpers on either side of the flow are mirrored too, as HBM-bound kernels of
libcnf (csrc/data_kernels.cu) on torch CUDA tensors: `down` / `up` (F:74-164), `preprocess_dataset_class`
(F:174-231), `preprocess_dataset_SR` (F:233-279), `de_logitify` (F:287-318), `instance_noise` (F:635-653),
`renew_noise` (F:660-676).  The reference maps them over a `tf.data.Dataset` element by element on the host; here the
"dataset" is a batched tensor [N,H,W,D] (an un-batched [H,W,D] element is accepted where the reference accepts one).
TFRecord parsing (F:26-65): `tfrecords.py`.
"""
import torch

from ._lib import Borrowed, check, lib, require_cuda, stream_ptr


def residual_block_plan(nb_channels_in, _which_dilations=(1, 2, 4), cardinality=4):
    """Channel bookkeeping of one dilated residual block: for every dilation d the branch gets
    nb_channels_in // d channels (F:579), split into `cardinality` groups of equal width (F:396-397)
    that read the FIRST nb_channels_in // d input channels (F:402); outputs are concatenated (F:590)."""
    branches, offset = [], 0
    for d in _which_dilations:
        nb = nb_channels_in // int(d)
        if cardinality == 1:
            groups, gin, gout = 1, nb_channels_in, nb
        else:
            assert not nb % cardinality
            groups, gin, gout = cardinality, nb // cardinality, nb // cardinality
        branches.append({'dilation': int(d), 'channels': nb, 'groups': groups, 'group_in': gin,
                         'group_out': gout, 'concat_offset': offset})
        offset += nb
    return {'branches': branches, 'concat_channels': offset}


def dilated_residual_block(y, nb_channels_in, nb_channels_out, _strides=(1, 1), _project_shortcut=False,
                           _which_dilations=[1, 2, 4], ksize=(4, 4), cardinality=4, ln=False, do=False,
                           ln_axis=-1, init='glorot_uniform', weights=None):
    """F:501-627 on a concrete torch CUDA tensor `y` [B,h,w,nb_channels_out]:
    LN(LReLU) -> Conv1x1(nb_channels_in) -> LN(LReLU) -> grouped dilated convs, concatenated over `_which_dilations`
    (branch d: nb_channels_in // d channels in `cardinality` groups that read the FIRST nb_channels_in // d channels)
    -> LN(LReLU) -> Conv1x1(nb_channels_out) -> + y.

    The reference is a Keras graph builder: every call creates NEW layers initialised with `init`.  Here the block runs
    through libcnf (`cnf_residual_block`, the launches a coupling layer issues for one of its blocks); `weights` is the
    block's parameter dict {'ln1.gamma', 'ln1.beta', 'pw1.kernel', 'pw1.bias', 'ln2.*', 'gc.d<d>.g<j>.kernel/bias',
    'ln3.*', 'pw2.kernel', 'pw2.bias'} (Keras shapes; the names `coupling_layer.get_weights()` uses after the 'rb<i>.'
    prefix).  weights=None initialises a fresh block with `init`, like the reference.  Pass `weights={}` to receive the
    freshly initialised parameters back in that dict.

    Strided / projected shortcuts, dropout and a bottleneck width different from the block width are never used by the
    reference model (M:1123-1130) and are not built; `ln_axis` must be -1 (the model's choice, M:1129)."""
    from .conv_cINN_make_model import coupling_layer
    if tuple(_strides) != (1, 1) or _project_shortcut or do:
        raise NotImplementedError("only the configuration conv_cINN_make_model uses is built "
                                  "(strides (1,1), identity shortcut, no dropout; M:1123-1130)")
    if nb_channels_in != nb_channels_out:
        raise NotImplementedError("identity shortcut needs nb_channels_in == nb_channels_out (F:614-623)")
    if ln_axis != -1:
        raise NotImplementedError("ln_axis must be -1 (whole-sample LayerNorm, F:350-360)")
    ks = ksize if isinstance(ksize, int) else ksize[0]
    if not isinstance(ksize, int) and ksize[0] != ksize[1]:
        raise NotImplementedError("square kernels only")
    y = require_cuda(y, "y")
    if y.dim() != 4 or y.shape[3] != nb_channels_out:
        raise ValueError(f"y: expected [B,h,w,{nb_channels_out}], got {tuple(y.shape)}")
    B, h, w, _ = y.shape
    # a checkerboard layer on a (2h, 2w, 1) tensor has the compressed shape (h, w, 2) and nk = num_kernels / 2 (M:420-423):
    # its residual blocks are exactly this block for any h, w
    layer = coupling_layer([2 * h, 2 * w, 1], 0, 1, cardinality, 2 * nb_channels_in, ks, init, LAYER_NORM=bool(ln),
                           which_dilations=list(_which_dilations), device=y.device)
    views = layer.weight_views()
    block_names = [k for k in views['A'] if k.startswith('rb0.')]
    if weights:
        missing = [k[4:] for k in block_names if k[4:] not in weights]
        if missing:
            raise ValueError(f"dilated_residual_block: weights lack {missing}")
        for k in block_names:
            for net in ('A', 'b'):
                src = torch.as_tensor(weights[k[4:]], dtype=torch.float32, device=y.device)
                views[net][k].copy_(src.reshape(views[net][k].shape))
    elif weights is not None:
        for k in block_names:
            weights[k[4:]] = views['A'][k].detach().cpu().numpy().copy()
    return layer.residual_block(y, 0, net="A")


# ---------------------------------------------------------------------------------------------------------------------
# data helpers (input / output side of the flow)
# ---------------------------------------------------------------------------------------------------------------------
def _batched(img, name):
    img = require_cuda(img, name)
    if img.dim() == 3:                       # F:91-96: an element without a batch dimension
        return img.unsqueeze(0), False
    if img.dim() != 4:
        raise ValueError(f"{name}: expected [H,W,D] or [B,H,W,D], got {tuple(img.shape)}")
    return img, True


def down(img, levels=1):
    """F:74-127: 2x2 average pool of an HxWxD image or a batch of them (trailing odd row / column cropped).
    `levels` > 1 nests the call (down(down(img)) = mean of means) inside one kernel."""
    x, batch = _batched(img, "img")
    B, H, W, D = x.shape
    out = torch.empty((B, H >> levels, W >> levels, D), dtype=torch.float32, device=x.device)
    br = Borrowed()
    check(lib.cnf_down(br(x), int(levels), br(out), stream_ptr()))
    return out if batch else out[0]


def up(img, levels=1):
    """F:129-164: 2x2 pixel repeat of an HxWxD image or a batch of them; `levels` nests the call."""
    x, batch = _batched(img, "img")
    B, H, W, D = x.shape
    out = torch.empty((B, H << levels, W << levels, D), dtype=torch.float32, device=x.device)
    br = Borrowed()
    check(lib.cnf_up(br(x), int(levels), br(out), stream_ptr()))
    return out if batch else out[0]


def preprocess_dataset_class(x_dataset, LOGITS=False, a=0.01):
    """F:174-231: identity unless LOGITS; then x -> logit(a + (1-a) b x) rescaled to [0,1]."""
    if not LOGITS:
        return x_dataset
    x = require_cuda(x_dataset, "x_dataset")
    out = torch.empty_like(x)
    br = Borrowed()
    check(lib.cnf_logit_scale(br(x), float(a), 0, br(out), stream_ptr()))
    return out


_SR_LEVELS = {'SR4,2': (1, 2), 'SR2,1': (0, 1)}


def preprocess_dataset_SR(x_dataset, model_type, RESIDUAL=True, levels=None):
    """F:233-279: (x, y) construction for super-resolution as ONE kernel: x = down^lx(hires),
    y = up^(ly-lx)(down^ly(hires)), x -= y if RESIDUAL, returns concat((x, y), -1).
    model_type 'SR4,2' -> (lx, ly) = (1, 2), 'SR2,1' -> (0, 1) as in the reference; `levels=(lx, ly)` generalises it
    (SURVEY config 4: 64x64 data with an 8x8 condition = (0, 3)).  Like the reference, any other model_type without
    `levels` is an error (there: UnboundLocalError on `xy_dataset`)."""
    if levels is None:
        if model_type not in _SR_LEVELS:
            raise UnboundLocalError(f"preprocess_dataset_SR: model_type {model_type!r} is not 'SR4,2' or 'SR2,1'")
        levels = _SR_LEVELS[model_type]
    lx, ly = (int(v) for v in levels)
    x, batch = _batched(x_dataset, "x_dataset")
    B, H, W, D = x.shape
    out = torch.empty((B, H >> lx, W >> lx, 2 * D), dtype=torch.float32, device=x.device)
    br = Borrowed()
    check(lib.cnf_sr_preprocess(br(x), lx, ly, 1 if RESIDUAL else 0, br(out), stream_ptr()))
    return out if batch else out[0]


def de_logitify(x, a=0.01):
    """F:287-318 on a CUDA tensor of samples (the reference takes numpy)."""
    x = require_cuda(x, "x")
    out = torch.empty_like(x)
    br = Borrowed()
    check(lib.cnf_logit_scale(br(x), float(a), 1, br(out), stream_ptr()))
    return out


class _NoiseState:
    """Philox (seed, offset) stream shared by instance_noise / renew_noise: every call consumes a fresh counter range,
    `manual_seed` restarts it (tf.random.set_seed analogue).  Ranks of a data-parallel job should seed with
    seed + rank so their shards draw independent noise."""
    seed = 0
    offset = 0


def manual_seed(seed):
    _NoiseState.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    _NoiseState.offset = 0


def _draw(x, alpha, out, seed, offset):
    n = out.numel()
    if seed is None:
        seed, offset = _NoiseState.seed, _NoiseState.offset
        _NoiseState.offset += (n + 3) // 4
    br = Borrowed()
    check(lib.cnf_instance_noise(br(x), float(alpha), int(seed), int(offset), br(out), stream_ptr()))
    return out


def instance_noise(x_element, alpha, seed=None, offset=0):
    """F:635-653: alpha x + (1 - alpha) N(0,1), noise generated on the device (Philox4x32-10 + Box-Muller)."""
    x = require_cuda(x_element, "x_element")
    return _draw(x, alpha, torch.empty_like(x), seed, offset)


def renew_noise(element, seed=None, offset=0):
    """F:660-676: a fresh N(0,1) tensor with the shape of `element` (its values are not read)."""
    x = require_cuda(element, "element")
    return _draw(None, 0.0, torch.empty_like(x), seed, offset)
