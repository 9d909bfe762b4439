"""Two devices in one process: the per-kernel attribute caches (opt-in shared memory, SM count) are
kept per device, so a second GPU used from the same process - here from its own thread, as cnf.h allows - launches with the
right limits and gives the same results.  Runs when at least two GPUs are visible."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(io_shape=[28, 28, 2], x_d=1, squeeze_factor_block_list=[0, 1, 0, 0], ResNeXt_block_list=[1, 1, 1, 1],
           num_kernels_list=[64, 64, 32, 32], cardinality_list=[8, 8, 4, 4])


def _run(dev, out):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import cFlow, Adam
    try:
        torch.cuda.set_device(dev)
        m = cFlow(**CFG, device=dev)
        m.randomize_weights(seed=3)
        x = torch.randn(6, 28, 28, 2, generator=torch.Generator().manual_seed(0)).to(dev)
        four = [float(t) for t in m.log_loss(x)]
        s = m(x, -1)
        m.compile(optimizer=Adam(3e-4))
        logs = m.train_step(x)
        torch.cuda.synchronize(dev)
        out[str(dev)] = (four, s.cpu(), float(logs['loss']))
    except Exception as e:      # surfaced by the main thread
        out[str(dev)] = e


def test_two_devices_one_process_threads():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    out = {}
    # device 1 FIRST and alone, then both concurrently: a process-wide cache would leave one of them without its attributes
    t = threading.Thread(target=_run, args=(torch.device("cuda:1"), out))
    t.start(); t.join()
    first = out["cuda:1"]
    assert not isinstance(first, Exception), first
    ts = [threading.Thread(target=_run, args=(torch.device(f"cuda:{i}"), out)) for i in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for k, v in out.items():
        assert not isinstance(v, Exception), (k, v)
    a, b = out["cuda:0"], out["cuda:1"]
    assert a[0] == pytest.approx(b[0], rel=1e-6)
    assert torch.allclose(a[1], b[1], rtol=1e-5, atol=1e-6)
    assert a[2] == pytest.approx(b[2], rel=1e-5)
    assert first[0] == pytest.approx(b[0], rel=1e-6)
