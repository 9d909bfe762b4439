"""Keras variable-name map of the reference model: creation order b-net before A-net (M:1106-1213),
per residual block LN1, 1x1, LN2, grouped convs (dilation-major), LN3, 1x1 (F:552-612)."""
import numpy as np
import pytest

from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow
from arl_conditional_normalizing_flows_b200.keras_interchange import (export_keras_npz, import_keras_npz,
                                                                      keras_weight_names)

CFG = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[2, 1],
           num_kernels_list=[16, 8], cardinality_list=[2, 2])


def test_name_map_order_and_counts():
    m = cFlow(**CFG, device="cpu")
    W = m.get_weights()
    table = keras_weight_names([{n: list(w) for n, w in lw.items()} for lw in W])
    names = [t[0] for t in table]
    assert len(set(names)) == len(names) == sum(len(w) for lw in W for w in lw.values())
    # the very first variables: b-net of coupling layer 0 (stem conv, then LN1 of block 0, then the 1x1 conv)
    assert table[0] == ('conv2d/kernel:0', 0, 'b', 'stem.kernel') and table[1] == ('conv2d/bias:0', 0, 'b', 'stem.bias')
    assert table[2] == ('layer_normalization/gamma:0', 0, 'b', 'rb0.ln1.gamma')
    assert table[4] == ('conv2d_1/kernel:0', 0, 'b', 'rb0.pw1.kernel')
    # grouped convs follow LN2 dilation-major, group-minor
    gc = [t[3] for t in table if t[1] == 0 and t[2] == 'b' and '.gc.' in t[3] and t[3].startswith('rb0') and t[3].endswith('kernel')]
    dil = m.coupling_layers[0].which_dilations
    assert gc == [f'rb0.gc.d{int(d)}.g{j}.kernel' for d in dil for j in range(2)]
    # the A-net of a layer comes after its whole b-net and ends with the unnamed tanh scale
    first_A = next(i for i, t in enumerate(table) if t[2] == 'A')
    assert all(t[2] == 'b' for t in table[:first_A]) and table[first_A][3] == 'stem.kernel'
    a0 = [t for t in table if t[1] == 0 and t[2] == 'A']
    assert a0[-1][0] == 'tanh_scaling_layer/Variable:0' and a0[-1][3] == 'tanh_scale'
    assert [t[0] for t in table if t[3] == 'tanh_scale'][1] == 'tanh_scaling_layer_1/Variable:0'
    # counters run across coupling layers; convs per net = stem + R (2 + n_dil * card) + head
    n_conv = sum(1 for n in names if n.endswith('kernel:0'))
    want = 0
    for layer in m.coupling_layers:
        info = layer._info
        want += 2 * (2 + info.R * (2 + sum(info.groups[i] for i in range(info.n_branches))))
    assert n_conv == want
    assert f'conv2d_{n_conv - 1}/kernel:0' in names and f'conv2d_{n_conv}/kernel:0' not in names
    shifted = keras_weight_names([{n: list(w) for n, w in lw.items()} for lw in W], offsets={'conv2d': 5})
    assert shifted[0][0] == 'conv2d_5/kernel:0'


def test_export_import_round_trip(tmp_path):
    m = cFlow(**CFG, device="cpu")
    rng = np.random.default_rng(0)
    W = m.get_weights()
    for lw in W:
        for net in lw.values():
            for k in net:
                net[k] = rng.standard_normal(np.shape(net[k])).astype(np.float32)
    m.set_weights(W)
    path = tmp_path / "keras_names.npz"
    names = export_keras_npz(m, str(path))
    z = np.load(path)
    assert sorted(z.files) == sorted(names)
    np.testing.assert_array_equal(z['conv2d/kernel:0'], W[0]['b']['stem.kernel'])
    assert z['tanh_scaling_layer/Variable:0'].shape == ()
    m2 = cFlow(**CFG, device="cpu")
    assert import_keras_npz(m2, str(path)) == len(names)
    for a, b in zip(m.get_weights(), m2.get_weights()):
        for net in ('A', 'b'):
            for k in a[net]:
                np.testing.assert_array_equal(a[net][k], b[net][k])
    bad = dict(z)
    bad['conv2d/kernel:0'] = bad['conv2d/kernel:0'][..., :1]
    np.savez(tmp_path / "bad.npz", **bad)
    with pytest.raises(ValueError, match="shape"):
        import_keras_npz(m2, str(tmp_path / "bad.npz"))
    del bad['conv2d/bias:0']
    np.savez(tmp_path / "bad2.npz", **bad)
    with pytest.raises(KeyError):
        import_keras_npz(m2, str(tmp_path / "bad2.npz"))
