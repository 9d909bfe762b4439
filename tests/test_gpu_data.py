"""Parity of the data-helper kernels (csrc/data_kernels.cu through the C-ABI) against oracle/data_np.py.

Tolerances: down / up / preprocess_dataset_SR are sums of <= 4 fp32 values per level in a fixed order and a multiply by
0.25 -> BIT-EXACT against the fp32 NumPy restatement; logit / de_logitify: 2e-6 absolute on values in [0, 1] (fp32 logf /
expf against the fp64 oracle); instance noise: the Philox integers are exact by construction, the Box-Muller normals are
checked at 2e-5 absolute against the fp64 restatement of the same stream.
"""
import numpy as np
import pytest
import torch

from oracle import data_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import arl_conditional_normalizing_flows_b200.conv_cINN_base_functions as F
    return F


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda:0")


@pytest.mark.parametrize("shape", [(3, 8, 12, 2), (1, 28, 28, 1), (2, 5, 7, 3), (2, 64, 64, 3)])
def test_down_up_bit_exact(F, shape):
    x = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
    np.testing.assert_array_equal(F.down(cuda(x)).cpu().numpy(), data_np.down(x))
    np.testing.assert_array_equal(F.up(cuda(x)).cpu().numpy(), data_np.up(x))
    np.testing.assert_array_equal(F.down(cuda(x[0])).cpu().numpy(), data_np.down(x[0]))     # un-batched element
    if shape[1] % 4 == 0 and shape[2] % 4 == 0:
        np.testing.assert_array_equal(F.down(cuda(x), levels=2).cpu().numpy(), data_np.down(data_np.down(x)))
        np.testing.assert_array_equal(F.up(cuda(x), levels=2).cpu().numpy(), data_np.up(data_np.up(x)))


@pytest.mark.parametrize("model_type,levels,residual", [('SR2,1', None, True), ('SR4,2', None, True),
                                                        ('SR2,1', None, False), (None, (0, 3), True),
                                                        (None, (1, 3), False)])
def test_preprocess_sr_bit_exact(F, model_type, levels, residual):
    h = np.random.default_rng(1).random((3, 32, 32, 3)).astype(np.float32)
    got = F.preprocess_dataset_SR(cuda(h), model_type, RESIDUAL=residual, levels=levels).cpu().numpy()
    want = data_np.preprocess_SR(h, model_type, RESIDUAL=residual, levels=levels)
    assert got.dtype == np.float32 and got.shape == want.shape
    np.testing.assert_array_equal(got, want)


def test_preprocess_sr_config4_shape(F):
    """64x64x3 images, 8x8 condition repeated to 64x64, residual x -> xy [B,64,64,6]."""
    h = np.random.default_rng(2).random((4, 64, 64, 3)).astype(np.float32)
    xy = F.preprocess_dataset_SR(cuda(h), None, levels=(0, 3)).cpu().numpy()
    assert xy.shape == (4, 64, 64, 6)
    np.testing.assert_allclose(xy[..., :3] + xy[..., 3:], h, atol=1e-6)
    y = xy[..., 3:].reshape(4, 8, 8, 8, 8, 3)
    assert np.all(y == y[:, :, :1, :, :1, :])                               # constant over every 8x8 block
    with pytest.raises(UnboundLocalError):
        F.preprocess_dataset_SR(cuda(h), 'class')
    with pytest.raises(AssertionError):
        F.preprocess_dataset_SR(cuda(h[:, :60]), None, levels=(0, 3))


def test_logit_and_de_logitify(F):
    x = np.random.default_rng(3).random((5, 28, 28, 1)).astype(np.float32)
    x.flat[:3] = (0.0, 1.0, 0.5)
    xt = cuda(x)
    assert F.preprocess_dataset_class(xt) is xt                               # LOGITS=False: untouched (F:207-209)
    s = F.preprocess_dataset_class(xt, LOGITS=True).cpu().numpy()
    np.testing.assert_allclose(s, data_np.preprocess_class_logits(x), rtol=0, atol=2e-6)
    back = F.de_logitify(cuda(s)).cpu().numpy()
    np.testing.assert_allclose(back, data_np.de_logitify(s.astype(np.float64)), rtol=0, atol=2e-6)
    np.testing.assert_allclose(back, x, rtol=0, atol=2e-5)
    s2 = F.preprocess_dataset_class(xt, LOGITS=True, a=0.05).cpu().numpy()
    np.testing.assert_allclose(s2, data_np.preprocess_class_logits(x, a=0.05), rtol=0, atol=2e-6)


@pytest.mark.parametrize("n_shape", [(4, 28, 28, 2), (1, 3, 3, 1), (2, 7, 5, 3)])
def test_instance_noise_matches_philox_oracle(F, n_shape):
    x = np.random.default_rng(4).standard_normal(n_shape).astype(np.float32)
    n = x.size
    for alpha in (0.0, 0.25, 1.0):
        got = F.instance_noise(cuda(x), alpha, seed=0x1234567890ABCDEF, offset=11).cpu().numpy()
        z = data_np.philox_normal(n, 0x1234567890ABCDEF, offset=11).reshape(n_shape)
        np.testing.assert_allclose(got, data_np.instance_noise(x, alpha, z), rtol=0, atol=2e-5)
    fresh = F.renew_noise(cuda(x), seed=5, offset=0).cpu().numpy()
    np.testing.assert_allclose(fresh, data_np.philox_normal(n, 5).reshape(n_shape), rtol=0, atol=2e-5)
    np.testing.assert_array_equal(F.instance_noise(cuda(x), 1.0, seed=1).cpu().numpy(), x)   # alpha = 1: x * 1 + 0 * z


def test_noise_stream_state_and_moments(F):
    x = torch.zeros(256, 28, 28, 2, device="cuda:0")
    F.manual_seed(7)
    a = F.renew_noise(x)
    b = F.renew_noise(x)
    F.manual_seed(7)
    a2 = F.renew_noise(x)
    assert torch.equal(a, a2) and not torch.equal(a, b)                       # consecutive calls use fresh counters
    z = torch.cat([a.flatten(), b.flatten()]).double()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3
    assert abs(float((z ** 4).mean()) - 3.0) < 0.03
    # the two calls continue ONE stream: b starts where a ended
    both = data_np.philox_normal(2 * x.numel(), 7)
    np.testing.assert_allclose(b.flatten()[:1000].cpu().numpy(), both[x.numel():x.numel() + 1000], atol=2e-5)


def test_errors(F):
    with pytest.raises(RuntimeError):
        F.down(torch.zeros(1, 4, 4, 1))                                         # CPU tensor: no fallback
    with pytest.raises(TypeError):
        F.up(torch.zeros(1, 4, 4, 1, device="cuda:0", dtype=torch.float64))
    with pytest.raises(ValueError):
        F.down(torch.zeros(4, 4, device="cuda:0"))
