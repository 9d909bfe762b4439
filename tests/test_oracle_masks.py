"""Pins the oracle's index permutations with the invariants the reference code implies
(SURVEY §4 items 1-2).  Bit-exact (integer/index work)."""
import numpy as np
import pytest

from oracle import masks_np as M
from oracle.planner import plan_flow


@pytest.mark.parametrize("D", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("m", [0, 1, 2, 3])
def test_mask_roundtrip_and_partition(m, D):
    rng = np.random.default_rng(D * 10 + m)
    u = rng.standard_normal((3, 6, 4, D)).astype(np.float32) + 5.0   # finite, no -0.0 (Q8)
    mc = {0: 1, 1: 0, 2: 3, 3: 2}[m]
    if D == 1 and m == 3:
        pytest.skip("mask 3 of a single channel is empty")
    # decompress(mask(u, m, True), m) == mask(u, m, False)   (M:500-761 vs M:763-1073)
    np.testing.assert_array_equal(M.decompress_mask(M.mask(u, m, True), m, u.shape), M.mask(u, m, False))
    # the two complementary masks partition u
    np.testing.assert_array_equal(M.mask(u, m, False) + M.mask(u, mc, False), u)


def closed_form_compressed(u, m):
    """The closed form SURVEY §8a A2 states (and the CUDA kernels use)."""
    D = u.shape[3]
    if m in (0, 1):
        a = 0 if m == 0 else 1
        return np.concatenate([u[:, 0::2, a::2, :], u[:, 1::2, (1 - a)::2, :]], -1)
    return u[..., (m - 2)::2]


@pytest.mark.parametrize("m", [0, 1, 2, 3])
def test_compressed_closed_form(m):
    u = np.arange(2 * 4 * 6 * 3, dtype=np.float32).reshape(2, 4, 6, 3)
    np.testing.assert_array_equal(M.mask(u, m, True), closed_form_compressed(u, m))


def test_space_to_depth_order():
    u = np.arange(1 * 4 * 4 * 3, dtype=np.float32).reshape(1, 4, 4, 3)
    v = M.space_to_depth2(u)
    for i in range(2):
        for j in range(2):
            for dy in range(2):
                for dx in range(2):
                    for c in range(3):
                        assert v[0, i, j, (dy * 2 + dx) * 3 + c] == u[0, 2 * i + dy, 2 * j + dx, c]
    np.testing.assert_array_equal(M.depth_to_space2(v), u)


def run_identity_flow(x, sq):
    """cFlow.call(+1) with every coupling replaced by identity (M:1743-1770)."""
    plan = plan_flow(list(x.shape[1:]), 1, sq, [1] * len(sq), [16] * len(sq), [2] * len(sq))
    sf = [L for L in plan['layers'] if L['type'] != 'coupling']
    uv, zy = x, None
    trace = []
    for L in plan['layers']:
        if L['type'] == 'squeeze':
            uv, zy = M.squeeze_forward(uv, zy)
        elif L['type'] == 'factor':
            uv, zy = M.factor_forward(uv, zy)
        else:
            trace.append(uv.copy())
    if not sf:
        return uv, trace
    zy = np.concatenate([zy, uv], 3)
    vu = None
    for L in reversed(sf):
        if L['type'] == 'squeeze':
            vu, zy = M.squeeze_backward(vu, zy)
        else:
            vu, zy = M.factor_backward(vu, zy, L['num_prev_factors'])
    return vu, trace


@pytest.mark.parametrize("sq", [[0, 1, 0, 0], [1, 1], [0, 1, 1, 0, 1], [1, 1, 1], [0, 0]])
def test_reassembly_is_identity(sq):
    H = W = 2 ** (sum(sq) + 1)
    x = np.arange(2 * H * W * 3, dtype=np.float32).reshape(2, H, W, 3)
    out, _ = run_identity_flow(x, sq)
    np.testing.assert_array_equal(out, x)


@pytest.mark.parametrize("sq", [[0, 1, 0, 0], [1, 1, 0], [0, 1, 1, 0]])
def test_active_tensor_is_a_strided_view(sq):
    """the active tensor at level L is rows h = 2^L-1 (mod 2^L) of the original buffer,
    each row reshaped (W/2^L, 2^L*D).  This is the addressing the CUDA path folds in."""
    H = W = 2 ** (sum(sq) + 1)
    D = 3
    x = np.arange(2 * H * W * D, dtype=np.float32).reshape(2, H, W, D)
    _, trace = run_identity_flow(x, sq)
    level = 0
    k = 0
    for i, s in enumerate(sq):
        for _ in range(4):
            act = trace[k]
            k += 1
            S = 2 ** level
            view = x[:, S - 1::S, :, :].reshape(2, H // S, W // S, S * D)
            np.testing.assert_array_equal(act, view)
        level += s
