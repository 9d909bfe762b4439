"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/cnf.h
declares, the C planner agrees with the oracle planner, the parameter layout round-trips, and the
error classes map to the reference's exception types.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle.planner import plan_flow, plan_coupling
from oracle.weights import init_weights, param_specs

pkg = pytest.importorskip("arl_conditional_normalizing_flows_b200")
from arl_conditional_normalizing_flows_b200 import _lib  # noqa: E402
from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import (  # noqa: E402
    cFlow, coupling_layer, squeeze_layer, factor_out_zy_layer, Layer)
from arl_conditional_normalizing_flows_b200.TOYcINN_make_model import cINN_affine  # noqa: E402

CFG2 = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
            num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])


def test_every_declared_symbol_is_exported(repo_root):
    declared = set()
    for name in ("cnf.h", "cnf_measure.h"):     # every header under include/ that declares entry points
        hdr = open(os.path.join(repo_root, "include", name)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        hdr = re.sub(r"#ifdef CNF_DEBUG.*?#endif", "", hdr, flags=re.S)      # debug-build-only hooks
        declared |= set(re.findall(r"\b(cnf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in cnf.h but not exported by libcnf.so"
    assert set(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ declared
    assert lib.cnf_version() == 100


@pytest.mark.parametrize("cfg", [
    CFG2,
    dict(io_shape=[32, 32, 4], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
         num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4]),
    dict(io_shape=[64, 64, 6], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
         num_kernels_list=[64, 64, 32, 32], cardinality_list=[4, 4, 2, 2]),
    dict(io_shape=[128, 128, 4], x_d=3, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[3] * 4,
         num_kernels_list=[64, 64, 32, 32], cardinality_list=[2, 2, 2, 2]),
    dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 1], ResNeXt_block_list=[2, 1],
         num_kernels_list=[16, 8], cardinality_list=[2, 2]),
])
def test_c_planner_matches_oracle_planner(cfg):
    m = cFlow(**cfg, device="cpu")
    p = plan_flow(cfg['io_shape'], cfg['x_d'], cfg['squeeze_factor_block_list'], cfg['ResNeXt_block_list'],
                  cfg['num_kernels_list'], cfg['cardinality_list'])
    assert list(m.scale_list) == p['scale_list']
    assert list(m.num_prev_factors_list) == p['num_prev_factors_list']
    assert m.io_shape_list.tolist() == p['io_shape_list']
    for a, b in zip(m.dilations_list, p['dilations_list']):
        assert a['checkerboard'] == [int(d) for d in b['checkerboard']]
        assert a['channelwise'] == [int(d) for d in b['channelwise']]
    kinds = {'coupling': coupling_layer, 'squeeze': squeeze_layer, 'factor': factor_out_zy_layer}
    assert [type(l) for l in m.layers_list] == [kinds[L['type']] for L in p['layers']]
    assert len(m.squeeze_factor_layers_list) == 2 * sum(cfg['squeeze_factor_block_list'])
    oc = [L for L in p['layers'] if L['type'] == 'coupling']
    for layer, L in zip(m.coupling_layers, oc):
        i = layer._info
        assert (i.h, i.w, i.c1, i.c2, i.nk, i.cat, i.mask, i.mask_complement) == \
               (L['h'], L['w'], L['c1'], L['c2'], L['nk'], L['cat'], L['mask'], L['mask_complement'])
        assert [i.dilation[j] for j in range(i.n_branches)] == L['dilations']
        assert [i.group_out[j] for j in range(i.n_branches)] == [b['group_width'] for b in L['branches']]
        # named parameters: same names and Keras shapes as the oracle's weight sets
        L2 = dict(L, ln=True)
        want = [(n, tuple(s)) for n, s, _ in param_specs(L2)] + [('tanh_scale', (1,))]
        got = [(n, s) for n, _, s, _ in layer._entries]
        assert sorted(got) == sorted(want)


def test_cfg2_parameter_count_matches_survey():
    m = cFlow(**CFG2, device="cpu")
    # 0.75 M conv params + 12.36 M LN params at config 2
    n = m.count_params()
    assert 13.0e6 < n < 13.3e6
    ln = sum(v.numel() for l in m.coupling_layers for d in l.weight_views().values()
             for k, v in d.items() if 'gamma' in k or 'beta' in k)
    assert abs(ln - 12.36e6) < 0.02e6


def test_weights_roundtrip_and_init_state():
    cfg = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
               num_kernels_list=[16, 8], cardinality_list=[2, 2])
    m = cFlow(**cfg, device="cpu")
    w0 = m.get_weights()
    for layer in w0:
        for net in ('A', 'b'):
            for k, v in layer[net].items():
                if k.endswith('gamma') or k == 'tanh_scale':
                    assert np.all(v == 1)
                elif k.endswith('bias') or k.endswith('beta'):
                    assert np.all(v == 0)
                else:   # Orthogonal(0.1): columns (or rows) orthonormal times gain
                    mat = v.reshape(-1, v.shape[-1]).astype(np.float64)
                    g = mat.T @ mat if mat.shape[0] >= mat.shape[1] else mat @ mat.T
                    np.testing.assert_allclose(g, 0.01 * np.eye(g.shape[0]), atol=1e-6)
    p = plan_flow(cfg['io_shape'], cfg['x_d'], cfg['squeeze_factor_block_list'], cfg['ResNeXt_block_list'],
                  cfg['num_kernels_list'], cfg['cardinality_list'])
    W = init_weights(p, 'rand', seed=4)
    m.set_weights(W)
    back = m.get_weights()
    for a, b in zip(W, back):
        for net in ('A', 'b'):
            assert set(a[net]) == set(b[net])
            for k in a[net]:
                np.testing.assert_array_equal(np.asarray(a[net][k]).reshape(-1), b[net][k].reshape(-1))
    with pytest.raises(ValueError):
        bad = [dict(A=dict(w['A']), b=dict(w['b'])) for w in W]
        bad[0]['A'].pop('stem.bias')
        m.set_weights(bad)


def test_reference_asserts_map_to_assertion_error():
    base = dict(io_shape=[8, 8, 2], x_d=1, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
                num_kernels_list=[16], cardinality_list=[2], device="cpu")
    with pytest.raises(AssertionError, match="spatial dimensions divisible by 2"):
        cFlow(**dict(base, io_shape=[7, 8, 2]))
    with pytest.raises(AssertionError, match="number of kernels"):
        cFlow(**dict(base, num_kernels_list=[15]))
    with pytest.raises(AssertionError, match="cardinality"):
        cFlow(**dict(base, cardinality_list=[3]))
    with pytest.raises(AssertionError, match="0 and 1"):
        cFlow(**dict(base, squeeze_factor_block_list=[2]))
    with pytest.raises(AssertionError, match="same length"):
        cFlow(**dict(base, ResNeXt_block_list=[1, 1]))
    with pytest.raises(AssertionError, match="cumulative scale"):
        cFlow(**dict(base, io_shape=[4, 4, 2], squeeze_factor_block_list=[1, 1, 0], ResNeXt_block_list=[1] * 3,
                     num_kernels_list=[16] * 3, cardinality_list=[2] * 3))
    with pytest.raises(AssertionError):      # Q6: (nk/2 // d) % cardinality on checkerboard layers (F:396)
        cFlow(io_shape=[64, 64, 6], x_d=3, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
              num_kernels_list=[64], cardinality_list=[8], device="cpu")
    with pytest.raises(AttributeError, match="dilations_list"):   # M:1553: DILATIONS=False cannot build
        cFlow(**dict(base, DILATIONS=False))
    with pytest.raises(NotImplementedError):
        Layer().forward_and_Jacobian(None, 0, None)
    with pytest.raises(NotImplementedError):
        Layer().backward(None, None)


def test_cpu_tensors_fail_loudly():
    m = cFlow(io_shape=[8, 8, 2], x_d=1, squeeze_factor_block_list=[0], ResNeXt_block_list=[1],
              num_kernels_list=[16], cardinality_list=[2], device="cpu")
    x = torch.zeros(2, 8, 8, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.log_loss(x)
    # the C side refuses a CPU DLPack tensor as well (device class -> RuntimeError)
    br = _lib.Borrowed()
    rc = _lib.lib.cnf_space_to_depth(br(x), br(torch.zeros(2, 4, 4, 8)), None)
    assert rc == _lib.CNF_ERR_DEVICE
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        _lib.check(rc)


def test_toy_model_host_logic():
    m = cINN_affine(3, 2, 12, 16, 2, None, device="cpu", seed=0)
    assert sorted(m.mask_indices[:6]) == list(range(6)) and sorted(m.mask_indices[6:]) == list(range(6, 12))
    assert list(m.dims_u1) == [1, 1, 1, 2, 2, 2] * 2 and list(m.dims_u2) == [2, 2, 2, 1, 1, 1] * 2
    w = m.coupling_layers_list[3].get_weights()
    assert len(w) == 4 * (2 + 2) and w[0].shape == (2, 16) and w[-2].shape == (16, 1)
    m2 = cINN_affine(3, 2, 12, 16, 2, None, mask_indices=list(range(12)), device="cpu")
    assert list(m2.mask_indices) == list(range(12))
    np.testing.assert_array_equal(m2.masks_1[3], np.eye(3, dtype=np.float32)[[0, 1]])
    m2._pack()
    assert m2.params.numel() == _lib.lib.cnf_toy_param_count(12, 16, 2)
