"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star):
  * squeeze / masking / permutation indexing and the pass-through half of a coupling layer: BIT-EXACT;
  * s/t nets, per-sample log-det, log-likelihood terms, samples from a fixed latent: 1e-4 relative
    (fp32 kernels vs the fp64 oracle; `rel` below is max|a-b| / max|b|, per-sample quantities are also
    checked element-wise with rtol 1e-4 and a small absolute floor).
"""
import numpy as np
import pytest
import torch

from oracle import masks_np
from oracle.flow_torch import FlowOracle
from oracle.planner import plan_coupling
from oracle.weights import init_weights, synth_inputs

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def assert_close_elementwise(a, b, rtol=RTOL, what=""):
    """element-wise gate next to the max-norm one: |a - b| <= rtol |b| + rtol rms(b) for EVERY element (the absolute floor
    is the tensor's rms, not its maximum, so small entries are held to a small error too)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * float(np.sqrt(np.mean(b * b))), err_msg=what)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def mk(cfg, kind='rand', seed=1, dev="cuda:0", dtype=torch.float64):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
    m = cFlow(**cfg, device=dev)
    o = FlowOracle(**cfg, dtype=dtype)
    W = init_weights(o.plan, kind, seed=seed)
    o.set_weights(W)
    m.set_weights(W)
    return m, o


TINY = dict(io_shape=[4, 4, 2], x_d=1, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
            num_kernels_list=[8], cardinality_list=[2])
SMALL = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
             num_kernels_list=[16, 8], cardinality_list=[2, 2])
MID = dict(io_shape=[16, 16, 4], x_d=3, squeeze_factor_block_list=[0, 1, 1], ResNeXt_block_list=[1, 2, 1],
           num_kernels_list=[32, 32, 16], cardinality_list=[4, 2, 2])
CFG2 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
CFG3 = dict(io_shape=[32, 32, 4], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])
# BASELINE configs 4 and 5 in their "light" variants (the reference's default lists violate
# `nb_channels % cardinality` on the checkerboard layers once dilation 8 appears, quirk Q6)
CFG4 = dict(io_shape=[64, 64, 6], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[4, 4, 2, 2])
CFG5 = dict(io_shape=[128, 128, 4], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[2, 2, 2, 2])


# ---------------------------------------------------------------------------------------------------
# index permutations: bit-exact
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("m", [0, 1, 2, 3])
def test_mask_and_decompress_bit_exact(dev, m, D):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
    if D == 1 and m == 3:
        pytest.skip("mask 3 of one channel is empty")
    L = coupling_layer([6, 8, D], 0, 1, 2, 8, 3, None, LAYER_NORM=True, which_dilations=[1], device=dev)
    rng = np.random.default_rng(7)
    u = rng.standard_normal((3, 6, 8, D)).astype(np.float32)
    ut = torch.from_numpy(u).to(dev)
    for compress in (False, True):
        got = L.mask(ut, m, compress).cpu().numpy()
        want = masks_np.mask(u, m, compress)
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
    uc = masks_np.mask(u, m, True)
    got = L.decompress_mask(torch.from_numpy(np.ascontiguousarray(uc)).to(dev), m, u.shape).cpu().numpy()
    np.testing.assert_array_equal(got, masks_np.decompress_mask(uc, m, u.shape))


def test_squeeze_and_factor_layers_bit_exact(dev):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import squeeze_layer, factor_out_zy_layer
    rng = np.random.default_rng(3)
    u = rng.standard_normal((2, 8, 12, 3)).astype(np.float32)
    zy = rng.standard_normal((2, 8, 12, 5)).astype(np.float32)
    sq, fa = squeeze_layer(), factor_out_zy_layer(1)
    v, s, z2 = sq.forward_and_Jacobian(torch.from_numpy(u).to(dev), 0.5, torch.from_numpy(zy).to(dev))
    wv, wz = masks_np.squeeze_forward(u, zy)
    np.testing.assert_array_equal(v.cpu().numpy(), wv)
    np.testing.assert_array_equal(z2.cpu().numpy(), wz)
    assert s == 0.5
    ub, zb = sq.backward(v, z2)
    np.testing.assert_array_equal(ub.cpu().numpy(), u)
    np.testing.assert_array_equal(zb.cpu().numpy(), zy)
    v3, _, z3 = fa.forward_and_Jacobian(v, 0, z2)
    wv3, wz3 = masks_np.factor_forward(wv, wz)
    np.testing.assert_array_equal(v3.cpu().numpy(), wv3)
    np.testing.assert_array_equal(z3.cpu().numpy(), wz3)
    u4, z4 = fa.backward(v3, z3)
    wu4, wz4 = masks_np.factor_backward(wv3, wz3, 1)
    np.testing.assert_array_equal(u4.cpu().numpy(), wu4)
    np.testing.assert_array_equal(z4.cpu().numpy(), wz4)
    with pytest.raises(AssertionError):
        sq.forward_and_Jacobian(torch.zeros(1, 3, 4, 2, device=dev), 0, None)
    with pytest.raises(AssertionError):
        sq.backward(torch.zeros(1, 2, 2, 6, device=dev), None)


# ---------------------------------------------------------------------------------------------------
# one coupling layer: s/t nets, forward, inverse
# ---------------------------------------------------------------------------------------------------
LAYER_CASES = [
    # in_shape, mask, R, card, num_kernels, dilations, LN
    ([8, 8, 3], 0, 2, 2, 16, [1, 2], True),
    ([8, 8, 3], 1, 1, 2, 16, [1], True),
    ([8, 8, 3], 2, 2, 2, 16, [1, 2], True),       # odd depth: c1=2, c2=1
    ([8, 8, 3], 3, 1, 2, 16, [1, 2], True),       # odd depth: c1=1, c2=2
    ([12, 8, 4], 2, 1, 4, 32, [1, 2], False),     # no LayerNorm
    ([28, 28, 2], 2, 3, 8, 64, [1, 2, 4], True),  # config-2 channel layer
    ([28, 28, 2], 0, 3, 8, 64, [1, 2, 4], True),  # config-2 checkerboard layer (nk 32, groups of 4/2/1)
    ([14, 14, 4], 3, 3, 4, 32, [1, 2], True),
    ([16, 16, 4], 2, 1, 1, 16, [1, 2], True),     # cardinality 1: plain conv over all channels (F:389-395)
    ([64, 64, 6], 0, 1, 4, 64, [1, 2, 4, 8], True),   # tiled grouped conv, wide head (c2 = 12)
    # shapes the activation-resident kernel (fused_kernels.cu) covers: config 2 / config 3 below the first level
    ([28, 28, 2], 1, 3, 8, 64, [1, 2, 4], True),  # (14,14,4) nk 32, groups of 4/2/1, 7-row segments
    ([14, 14, 4], 0, 3, 4, 32, [1, 2], True),     # (7,7,8) nk 16, c2 = 8
    ([14, 14, 4], 2, 3, 4, 32, [1, 2], True),     # (14,14,2) nk 32, groups of 8/4
    ([32, 32, 4], 0, 3, 8, 64, [1, 2, 4], True),  # (16,16,8) nk 32: one 512-thread CTA per SM, 8-row segments
    ([16, 16, 8], 3, 2, 4, 32, [1, 2], True),     # (16,16,4) nk 32
    ([16, 16, 8], 1, 2, 4, 32, [1, 2], True),     # (8,8,16) nk 16: c1 = 16, c2 = 16 -> layer-per-kernel path
    ([16, 16, 6], 2, 2, 4, 32, [1, 2, 4], True),  # (16,16,3) nk 32, groups 8/4/2
    ([16, 16, 5], 2, 1, 2, 16, [1, 2], True),     # odd depth through the resident kernel: c1 = 3, c2 = 2
    # groups of 16 / 32 channels: the tcgen05 implicit-GEMM grouped conv (gconv_tc.cuh), configs 4 / 5 layer shapes
    ([64, 64, 6], 2, 1, 4, 64, [1, 2, 4, 8], True),   # config-4 (light) channel layer: groups of 16 / 8 / 4 / 2
    ([32, 32, 4], 2, 2, 2, 64, [1, 2], True),         # groups of 32 (dil 1) and 16 (dil 2): config-5 (light) widths
    ([20, 12, 4], 2, 1, 2, 32, [1, 2], True),         # ragged 16 x 8 tiles (20 rows, 12 columns), groups of 16
    ([64, 64, 2], 0, 1, 2, 64, [1, 2, 4, 8, 16], True),   # config-5 (light) checkerboard layer: (32,32,4), cat = 62 (64-bit stores)
]


@pytest.mark.parametrize("fuse", [1, 0], ids=["resident", "per-kernel"])
@pytest.mark.parametrize("case", LAYER_CASES, ids=[f"{c[0]}m{c[1]}" for c in LAYER_CASES])
def test_coupling_layer_vs_oracle(dev, case, fuse):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
    from oracle.flow_torch import st_net, _t
    shape, m, R, card, nk, dil, ln = case
    L = plan_coupling(shape, m, R, card, nk, 3, dil)
    L['ln'] = ln
    W = init_weights({'layers': [L]}, 'rand', seed=11, ln=ln)[0]
    layer = coupling_layer(shape, m, R, card, nk, 3, None, LAYER_NORM=ln, which_dilations=dil, device=dev)
    layer.set_weights(W)
    layer.set_fusion(fuse)      # activation-resident launch (where the layer fits it) / layer-per-kernel path
    B = 3
    rng = np.random.default_rng(5)
    u = rng.standard_normal((B, *shape)).astype(np.float32)

    # oracle in fp64
    u1c = masks_np.mask(u.astype(np.float64), m, True)
    PA = {k: _t(v, torch.float64) for k, v in W['A'].items()}
    Pb = {k: _t(v, torch.float64) for k, v in W['b'].items()}
    with torch.no_grad():
        A = st_net(_t(u1c, torch.float64), PA, L, True).numpy()
        b = st_net(_t(u1c, torch.float64), Pb, L, False).numpy()
    mc = L['mask_complement']
    u2c = masks_np.mask(u.astype(np.float64), mc, True)
    v_want = masks_np.mask(u.astype(np.float64), m, False) + masks_np.decompress_mask(np.exp(A) * u2c + b, mc, u.shape)
    ld_want = A.sum(axis=(1, 2, 3))

    # A_wrapper / b_wrapper on the compressed input
    ut = torch.from_numpy(u).to(dev)
    u1c_t = layer.mask(ut, m, True)
    A_got = layer.A_wrapper(u1c_t).cpu().numpy()
    b_got = layer.b_wrapper(u1c_t).cpu().numpy()
    assert rel(A_got, A) < RTOL, rel(A_got, A)
    assert rel(b_got, b) < RTOL, rel(b_got, b)

    # forward_and_Jacobian
    v, s, zy = layer.forward_and_Jacobian(ut, 0.25, None)
    v = v.cpu().numpy()
    assert zy is None
    keep = masks_np.mask(np.ones_like(u), m, False) == 1
    assert np.array_equal(v[keep], u[keep]), "pass-through half must be bit-exact"
    assert rel(v, v_want) < RTOL, rel(v, v_want)
    ld = layer.last_logdet_per_sample.cpu().numpy()
    np.testing.assert_allclose(ld, ld_want, rtol=RTOL, atol=RTOL * np.abs(ld_want).mean())
    np.testing.assert_allclose(float(s), 0.25 + ld_want.mean(), rtol=RTOL, atol=1e-5)

    # backward inverts forward, and matches the oracle's inverse law on fresh input
    ub, _ = layer.backward(torch.from_numpy(v).to(dev), None)
    assert rel(ub.cpu().numpy(), u) < RTOL
    u_inv_want = masks_np.mask(u.astype(np.float64), m, False) + \
        masks_np.decompress_mask((u2c - b) / np.exp(A), mc, u.shape)
    u_inv, _ = layer.backward(ut, None)
    assert rel(u_inv.cpu().numpy(), u_inv_want) < RTOL

    with pytest.raises(ValueError):      # tf.ensure_shape (M:1276)
        layer.forward_and_Jacobian(torch.zeros(2, shape[0], shape[1] + 2, shape[2], device=dev), 0, None)


# ---------------------------------------------------------------------------------------------------
# the flow
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,cfg,B,kind", [
    ("tiny", TINY, 5, 'rand'), ("small", SMALL, 4, 'rand'), ("small-init", SMALL, 4, 'init'),
    ("mid", MID, 3, 'rand'), ("cfg2", CFG2, 2, 'rand'), ("cfg2-init", CFG2, 2, 'init'), ("cfg3", CFG3, 2, 'rand'),
    ("cfg4", CFG4, 2, 'rand'), ("cfg5", CFG5, 1, 'rand'),
])
def test_flow_forward_inverse_loss_vs_oracle(dev, name, cfg, B, kind):
    m, o = mk(cfg, kind, seed=2)
    H, W, D = cfg['io_shape']
    if name.startswith("cfg2"):
        x = synth_inputs('cfg2', B, seed=3)
    elif name in ("cfg3", "cfg4", "cfg5"):
        x = synth_inputs(name, B, seed=3)
    else:
        x = synth_inputs(f'noise:{H}x{W}x{D}', B, seed=3)
    xt = torch.from_numpy(x).to(dev)
    four_want, ps = o.log_loss(x.astype(np.float64))

    zy, ld = m(xt, 1)
    assert rel(zy.cpu().numpy(), ps['zy']) < RTOL
    assert_close_elementwise(zy.cpu().numpy(), ps['zy'], what="zy")
    ld_ps = m.last_logdet_per_sample.cpu().numpy()
    np.testing.assert_allclose(ld_ps, ps['logdet'], rtol=RTOL, atol=RTOL * np.abs(ps['logdet']).mean())
    np.testing.assert_allclose(float(ld), ps['logdet'].mean(), rtol=RTOL, atol=RTOL * np.abs(ps['logdet']).mean())
    assert torch.equal(xt, torch.from_numpy(x).to(dev)), "the public call must not mutate its input"

    four = [float(t) for t in m.log_loss(xt)]
    np.testing.assert_allclose(m.last_per_sample['ll_z'].cpu().numpy(), ps['ll_z'], rtol=RTOL)
    np.testing.assert_allclose(m.last_per_sample['ll_y'].cpu().numpy(), ps['ll_y'], rtol=RTOL,
                               atol=RTOL * np.abs(ps['ll_y']).mean())
    np.testing.assert_allclose(four[0], four_want[0], rtol=RTOL)
    np.testing.assert_allclose(four[1], four_want[1], rtol=RTOL)
    np.testing.assert_allclose(four[2], four_want[2], rtol=RTOL)
    np.testing.assert_allclose(four[3], four_want[3], rtol=RTOL, atol=RTOL * np.abs(ps['logdet']).mean())
    bpd = m.bits_per_dim(four[1], four[3])
    from oracle.flow_torch import bits_per_dim
    np.testing.assert_allclose(bpd, bits_per_dim(four_want[1], four_want[3], H, W, cfg['x_d']), rtol=RTOL, atol=1e-6)

    # sampling from a fixed latent
    z = synth_inputs(f'noise:{H}x{W}x{D}', B, seed=9)
    z[..., cfg['x_d']:] = x[..., cfg['x_d']:]
    xs_want = o.call(z.astype(np.float64), -1)
    xs = m(torch.from_numpy(z).to(dev), -1)
    assert rel(xs.cpu().numpy(), xs_want) < RTOL
    assert_close_elementwise(xs.cpu().numpy(), xs_want, what="samples")
    # default direction is -1 (M:1725)
    assert torch.equal(m(torch.from_numpy(z).to(dev)), xs)

    td = m.test_step(xt)
    assert set(td) == {'loss', 'z_loss', 'y_loss', 'detJ_loss'}
    np.testing.assert_allclose(td['loss'], four_want[0], rtol=RTOL)


@pytest.mark.parametrize("name,cfg,B", [("cfg2", CFG2, 33), ("cfg3", CFG3, 5), ("mid", MID, 4)])
def test_resident_path_equals_per_kernel_path(dev, name, cfg, B):
    """The activation-resident launch and the layer-per-kernel path are two implementations of the same layer: same
    zy / per-sample log-det / samples to fp32 rounding (summation orders differ), and identical pass-through halves."""
    m, _ = mk(cfg, 'rand', seed=7, dtype=torch.float32)
    H, W, D = cfg['io_shape']
    x = torch.from_numpy(synth_inputs(name if name.startswith('cfg') else f'noise:{H}x{W}x{D}', B, seed=4)).to(dev)
    out = {}
    for fuse in (1, 0):
        m.set_fusion(fuse)
        zy, _ = m(x, 1)
        out[fuse] = (zy.clone(), m.last_logdet_per_sample.clone(), m(x, -1).clone())
    for a, b_ in zip(out[1], out[0]):
        assert torch.isfinite(a).all()
        err = float((a - b_).abs().max() / b_.abs().max())
        assert err < 2e-5, err
    # sub-batches reproduce the full batch bit for bit on the resident path (one CTA per sample, fixed reduction order)
    m.set_fusion(1)
    zy_k, _ = m(x[:3].contiguous(), 1)
    assert torch.equal(zy_k, out[1][0][:3])


PATH_SETS = {   # CNF_PATH_* bits (include/cnf.h): which kernel serves each stage
    "per-kernel": 1, "pw-ffma": 1 | 2, "pw-gemm": 1 | 2 | 4, "gconv-branch": 1 | 8, "gconv2": 1 | 8 | 16,
    "gconv-generic": 1 | 8 | 16 | 32, "stem-gemm+head-tiles": 1 | 64 | 128, "all-general": 255,
}


@pytest.mark.parametrize("shape,m,card,nk,dil", [
    ([28, 28, 2], 2, 8, 64, [1, 2, 4]), ([28, 28, 2], 0, 8, 64, [1, 2, 4]), ([14, 14, 4], 3, 4, 32, [1, 2]),
    ([32, 32, 4], 2, 8, 64, [1, 2, 4]), ([64, 64, 6], 1, 4, 64, [1, 2, 4, 8]), ([12, 10, 3], 2, 2, 16, [1, 2]),
], ids=lambda v: str(v))
def test_every_kernel_family_gives_the_same_layer(dev, shape, m, card, nk, dil):
    """Each stage of a coupling layer has a fastest kernel and general ones for the shapes it does not cover.  Excluding
    kernel families (cnf_coupling_set_kernel_paths) must not change the layer: same v, log-det and inverse to fp32
    rounding on every path, and the default path agrees with the oracle elsewhere in this file."""
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
    L = plan_coupling(shape, m, 2, card, nk, 3, dil)
    W = init_weights({'layers': [L]}, 'rand', seed=13)[0]
    layer = coupling_layer(shape, m, 2, card, nk, 3, None, LAYER_NORM=True, which_dilations=dil, device=dev)
    layer.set_weights(W)
    g = torch.Generator(device=dev).manual_seed(5)
    u = torch.randn(5, *shape, device=dev, generator=g)
    layer.set_kernel_paths(0)
    v0, _, _ = layer.forward_and_Jacobian(u, 0.0, None)
    ld0 = layer.last_logdet_per_sample.clone()
    ui0, _ = layer.backward(u, None)
    assert torch.isfinite(v0).all() and torch.isfinite(ui0).all()
    for name, bits in PATH_SETS.items():
        layer.set_kernel_paths(bits)
        v, _, _ = layer.forward_and_Jacobian(u, 0.0, None)
        ld = layer.last_logdet_per_sample
        ui, _ = layer.backward(u, None)
        for a, b_ in ((v, v0), (ld, ld0), (ui, ui0)):
            err = float((a - b_).abs().max() / b_.abs().max())
            assert err < 2e-5, (name, err)
    with pytest.raises(AssertionError):
        layer.set_kernel_paths(256)


def test_flow_layerwise_equals_fused_call(dev):
    """Driving layers_list by hand exactly like cFlow.call (M:1743-1770) gives the fused in-place result."""
    m, _ = mk(MID, 'rand', seed=4)
    x = torch.from_numpy(synth_inputs('noise:16x16x4', 3, seed=1)).to(dev)
    uv, ld, zy = x, 0, None
    for layer in m.layers_list:
        uv, ld, zy = layer.forward_and_Jacobian(uv, ld, zy)
    zy = torch.cat([zy, uv], 3)
    vu = None
    for layer in reversed(m.squeeze_factor_layers_list):
        vu, zy = layer.backward(vu, zy)
    fused, ld_f = m(x, 1)
    assert torch.equal(vu, fused)
    np.testing.assert_allclose(float(ld), float(ld_f), rtol=1e-5)
    # and the mirror image for direction -1 (M:1774-1798)
    uv, zy = fused, None
    for layer in m.squeeze_factor_layers_list:
        uv, _, zy = layer.forward_and_Jacobian(uv, None, zy)
    vu = uv
    for layer in reversed(m.layers_list):
        vu, zy = layer.backward(vu, zy)
    assert torch.equal(vu, m(fused, -1))


def test_cfg2_full_batch_against_the_fp64_oracle(dev):
    """BASELINE config 2 at its FULL batch (256) against the fp64 oracle: tests/golden/cfg2_b256_digest.npz
    (oracle/make_golden.py) holds the four loss scalars, every per-sample log-det / ll_z / ll_y, every 7th element of zy
    and of the samples drawn from a fixed latent, and per-sample fp64 sums of both tensors."""
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg2_b256_digest.npz"))
    cfg = json.loads(str(g['config']))
    m, _ = mk(cfg, 'rand', seed=int(g['weights_seed']), dtype=torch.float32)
    B, stride = 256, int(g['stride'])
    x = synth_inputs('cfg2', B, seed=int(g['xy_seed']))
    four = [float(t) for t in m.log_loss(torch.from_numpy(x).to(dev))]
    np.testing.assert_allclose(four, g['loss4'], rtol=RTOL)
    ps = m.last_per_sample
    for k in ('logdet', 'll_z', 'll_y'):
        np.testing.assert_allclose(ps[k].cpu().numpy(), g[k], rtol=RTOL, atol=RTOL * np.abs(g[k]).mean(), err_msg=k)
    zy = ps['zy'].cpu().numpy().reshape(B, -1).astype(np.float64)
    assert rel(zy[:, ::stride], g['zy_strided']) < RTOL
    assert_close_elementwise(zy[:, ::stride], g['zy_strided'], what="zy")
    np.testing.assert_allclose(zy.sum(1), g['zy_sum'], rtol=RTOL, atol=RTOL * g['zy_abs'].mean())
    z = synth_inputs('noise:28x28x2', B, seed=int(g['z_seed']))
    z[..., 1:] = x[..., 1:]
    xs = m(torch.from_numpy(z).to(dev), -1).cpu().numpy().reshape(B, -1).astype(np.float64)
    assert rel(xs[:, ::stride], g['xs_strided']) < RTOL
    assert_close_elementwise(xs[:, ::stride], g['xs_strided'], what="samples")
    np.testing.assert_allclose(xs.sum(1), g['xs_sum'], rtol=RTOL, atol=RTOL * g['xs_abs'].mean())
    # the same on the layer-per-kernel path
    m.set_fusion(0)
    four0 = [float(t) for t in m.log_loss(torch.from_numpy(x).to(dev))]
    np.testing.assert_allclose(four0, g['loss4'], rtol=RTOL)
    np.testing.assert_allclose(m.last_per_sample['logdet'].cpu().numpy(), g['logdet'], rtol=RTOL,
                               atol=RTOL * np.abs(g['logdet']).mean())


def test_full_size_roundtrip_cfg2(dev):
    """BASELINE config 2 at its full batch: size-independent properties."""
    x = torch.from_numpy(synth_inputs('cfg2', 256, seed=6)).to(dev)
    # encode -> decode round trip: tight at the reference's initial state, conditioning-limited (the fp32
    # CPU oracle itself is ~1e-3 off) with the "trained-like" random weights
    for kind, tol in (('init', 1e-5), ('rand', 2e-2)):
        m, _ = mk(CFG2, kind, seed=5, dtype=torch.float32)
        zy, ld = m(x, 1)
        assert torch.isfinite(zy).all() and torch.isfinite(m.last_logdet_per_sample).all()
        xb = m(zy, -1)
        assert float((xb - x).abs().max() / x.abs().max()) < tol, kind
    # batch independence: any sub-batch gives the same per-sample results (no cross-sample op)
    zy_a, _ = m(x[:7].contiguous(), 1)
    ld_a = m.last_logdet_per_sample.clone()
    m(x, 1)
    assert float((zy_a - zy[:7]).abs().max()) < 1e-5
    assert float((ld_a - m.last_logdet_per_sample[:7]).abs().max() / ld_a.abs().max()) < 1e-5
    # loss decomposition (M:1840-1848)
    loss, zl, yl, dl = (float(t) for t in m.log_loss(x))
    np.testing.assert_allclose(loss, zl + yl + dl, rtol=1e-5)


def test_edge_cases(dev):
    m, o = mk(SMALL, 'rand', seed=2)
    # batch of one, and an empty batch
    x = synth_inputs('noise:8x8x3', 1, seed=3)
    zy, _ = m(torch.from_numpy(x).to(dev), 1)
    want, _, _ = o.call(x.astype(np.float64), 1)
    assert rel(zy.cpu().numpy(), want) < RTOL
    e = torch.zeros(0, 8, 8, 3, device=dev)
    zy, ld = m(e, 1)
    assert zy.shape == (0, 8, 8, 3)
    assert m(e, -1).shape == (0, 8, 8, 3)
    with pytest.raises(AssertionError):
        m.log_loss(e)
    with pytest.raises(ValueError):
        m(torch.zeros(2, 8, 8, 4, device=dev), 1)
    with pytest.raises(TypeError):
        m(torch.zeros(2, 8, 8, 3, device=dev, dtype=torch.float64), 1)
    assert m(torch.zeros(1, 8, 8, 3, device=dev), 0) is None      # M:1743/M:1774: neither branch
    # non-contiguous input is accepted (made compact on the host side)
    xx = torch.from_numpy(synth_inputs('noise:8x8x3', 4, seed=8)).to(dev)
    a, _ = m(xx[::2], 1)
    b_, _ = m(xx[::2].contiguous(), 1)
    assert torch.equal(a, b_)


# ---------------------------------------------------------------------------------------------------
# standalone fused coupling-law kernel
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [[8, 8, 3], [28, 28, 2], [6, 10, 5], [32, 32, 4], [12, 20, 8]])
@pytest.mark.parametrize("m", [0, 1, 2, 3])
def test_coupling_law_kernel(dev, shape, m):
    from arl_conditional_normalizing_flows_b200 import _lib
    rng = np.random.default_rng(m)
    B = 5
    u = rng.standard_normal((B, *shape))
    mc = m ^ 1
    u2c = masks_np.mask(u, mc, True)
    s = 0.5 * rng.standard_normal(u2c.shape)
    t = rng.standard_normal(u2c.shape)
    want_f = masks_np.mask(u, m, False) + masks_np.decompress_mask(np.exp(s) * u2c + t, mc, u.shape)
    want_i = masks_np.mask(u, m, False) + masks_np.decompress_mask((u2c - t) / np.exp(s), mc, u.shape)
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    ut, st_, tt = f32(u), f32(s), f32(t)
    for inverse, want in ((0, want_f), (1, want_i)):
        v = torch.empty_like(ut)
        ld = torch.empty(B + (-B) % 4, device=dev)[:B]
        br = _lib.Borrowed()
        _lib.check(_lib.lib.cnf_coupling_law(br(ut), br(st_), br(tt), m, inverse, br(v), br(ld), _lib.stream_ptr()))
        assert rel(v.cpu().numpy(), want) < 1e-5
        keep = masks_np.mask(np.ones_like(u), m, False) == 1
        assert np.array_equal(v.cpu().numpy()[keep], u.astype(np.float32)[keep])
        np.testing.assert_allclose(ld.cpu().numpy(), s.astype(np.float32).sum(axis=(1, 2, 3)), rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------------
# toy model (config 1)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("width,num_layers,n", [(32, 6, 24), (16, 2, 12), (8, 0, 6)])
def test_toy_vs_oracle(dev, width, num_layers, n):
    from arl_conditional_normalizing_flows_b200.TOYcINN_make_model import cINN_affine
    from oracle.toy import ToyOracle, toy_init_weights
    W = toy_init_weights(n, width, num_layers, seed=3, scale=1.0)   # scale 1.5 is chaotic even in the fp32 oracle
    order = list(np.random.default_rng(1).permutation(n))
    m = cINN_affine(3, 2, n, width, num_layers, None, mask_indices=order, device=dev)
    for j, cl in enumerate(m.coupling_layers_list):
        flat = []
        for net in ('b', 'A'):
            for Wm, bv in W[j][net]:
                flat += [Wm, bv]
        cl.set_weights(flat)
    o = ToyOracle(3, 2, n, W, mask_indices=order, dtype=np.float64)
    rng = np.random.default_rng(0)
    B = 1000
    xy = np.concatenate([rng.standard_normal((B, 2)), np.where(rng.uniform(size=(B, 1)) < 0.5, -1.0, 1.0)], 1)
    four_want, ps = o.log_loss(xy)
    xt = torch.from_numpy(xy.astype(np.float32)).to(dev)
    zy, ld = m(xt, -1)
    assert rel(zy.cpu().numpy(), ps['zy']) < RTOL
    np.testing.assert_allclose(ld.cpu().numpy(), ps['logdet'], rtol=RTOL, atol=RTOL * np.abs(ps['logdet']).mean())
    four = [float(t) for t in m.log_loss(xt)]
    np.testing.assert_allclose(four, four_want, rtol=RTOL, atol=1e-5)
    back, zero = m(zy, 1)
    assert zero == 0
    assert rel(back.cpu().numpy(), xy) < RTOL
    want_inv, _ = o.call(xy, 1)
    got_inv, _ = m(xt, 1)
    assert rel(got_inv.cpu().numpy(), want_inv) < RTOL


@pytest.mark.parametrize("shape,nk,card,dil,B", [([28, 28, 2], 64, 8, [1, 2, 4], 9), ([14, 14, 4], 32, 4, [1, 2], 9),
                                                 ([14, 14, 4], 32, 4, [1, 2], 70), ([32, 32, 4], 64, 8, [1, 2, 4], 5)])
def test_st_nets_are_batch_independent_bitwise(dev, shape, nk, card, dil, B):
    """The fused grouped-conv kernel packs several samples into one item (ragged last item, TMA tiles, two buffers); its
    LayerNorm statistics are reduced in a fixed per-sample order, so any sub-batch reproduces the full batch bit for bit."""
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import coupling_layer
    layer = coupling_layer(shape, 2, 2, card, nk, 3, None, LAYER_NORM=True, which_dilations=dil, device=dev)
    info = layer._info
    g = torch.Generator(device=dev).manual_seed(11)
    u = torch.randn(B, info.h, info.w, info.c1, device=dev, generator=g)
    A, b = layer.A_wrapper(u).clone(), layer.b_wrapper(u).clone()
    assert torch.isfinite(A).all() and torch.isfinite(b).all()
    for k in (1, 2, B - 1):
        Ak, bk = layer.A_wrapper(u[:k].contiguous()), layer.b_wrapper(u[:k].contiguous())
        assert torch.equal(Ak, A[:k]) and torch.equal(bk, b[:k]), k
    # and a second evaluation of the same batch is identical (no order-dependent accumulation)
    assert torch.equal(layer.A_wrapper(u), A)
