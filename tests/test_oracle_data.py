"""CPU checks of the data-helper oracle (oracle/data_np.py): algebraic identities of F:74-318 and the Random123
known-answer vectors for Philox4x32-10."""
import numpy as np

from oracle import data_np


def test_down_up_identities():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 8, 12, 2)).astype(np.float32)
    np.testing.assert_array_equal(data_np.down(data_np.up(x)), x)           # exact: mean of four equal values
    assert data_np.up(x).shape == (3, 16, 24, 2)
    assert data_np.down(x).shape == (3, 4, 6, 2)
    assert data_np.down(x[0]).shape == (4, 6, 2)                           # element without a batch axis (F:91-96)
    # odd trailing row / column is cropped (F:107-110)
    y = rng.standard_normal((1, 5, 7, 1)).astype(np.float32)
    np.testing.assert_array_equal(data_np.down(y), data_np.down(y[:, :4, :6]))
    np.testing.assert_allclose(data_np.down(x), x.reshape(3, 4, 2, 6, 2, 2).mean(axis=(2, 4)), rtol=1e-6, atol=1e-7)


def test_sr_preprocess_structure():
    rng = np.random.default_rng(1)
    h = rng.random((2, 16, 16, 3)).astype(np.float32)
    xy = data_np.preprocess_SR(h, 'SR2,1')
    assert xy.shape == (2, 16, 16, 6)
    np.testing.assert_allclose(xy[..., :3] + xy[..., 3:], h, rtol=0, atol=1e-6)     # residual + condition = image
    xy42 = data_np.preprocess_SR(h, 'SR4,2', RESIDUAL=False)
    assert xy42.shape == (2, 8, 8, 6)
    np.testing.assert_array_equal(xy42[..., :3], data_np.down(h))
    # the residual has zero mean over every condition block
    xy3 = data_np.preprocess_SR(h, levels=(0, 3))
    blk = xy3[..., :3].reshape(2, 2, 8, 2, 8, 3).mean(axis=(2, 4))
    assert np.abs(blk).max() < 1e-6


def test_logit_round_trip():
    x = np.linspace(0, 1, 101)
    s = data_np.preprocess_class_logits(x)
    assert s.min() == 0.0 and abs(s.max() - 1.0) < 1e-12 and np.all(np.diff(s) > 0)
    np.testing.assert_allclose(data_np.de_logitify(s), x, atol=1e-12)


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = data_np.philox4x32_10(np.array(ctr, dtype=np.uint64), key)
        assert tuple(int(v) for v in got) == want


def test_philox_normal_moments():
    z = data_np.philox_normal(200_000, seed=123)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs((z ** 3).mean()) < 0.03 and abs((z ** 4).mean() - 3.0) < 0.06
    # offset semantics: a stream started at counter q equals the tail of the stream started at 0
    np.testing.assert_array_equal(data_np.philox_normal(40, 7, offset=5), data_np.philox_normal(60, 7)[20:])
