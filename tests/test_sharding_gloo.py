"""N>1 path on CPU: world_size-2 gloo run of the batch-sharding logic.  Each rank evaluates the
oracle on its contiguous shard; the reduced scalars must equal the single-process full-batch loss."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.flow_torch import FlowOracle
from oracle.weights import init_weights, synth_inputs

pytest.importorskip("arl_conditional_normalizing_flows_b200")
from arl_conditional_normalizing_flows_b200.sharding import (allreduce_mean_gradients, global_loss,  # noqa: E402
                                                             shard_bounds)

CFG = dict(io_shape=[8, 8, 3], x_d=2, squeeze_factor_block_list=[1, 0], ResNeXt_block_list=[1, 1],
           num_kernels_list=[16, 8], cardinality_list=[2, 2])


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    o = FlowOracle(**CFG, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=3))
    x = synth_inputs('noise:8x8x3', B, seed=4).astype(np.float64)
    lo, hi = shard_bounds(B, rank, world)
    _, ps = o.log_loss(x[lo:hi])
    four = global_loss(torch.from_numpy(ps['ll_z']), torch.from_numpy(ps['ll_y']), torch.from_numpy(ps['logdet']))
    q.put((rank, four, (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 5])      # even and uneven shards
def test_two_rank_sharded_loss_equals_single_process(B):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = FlowOracle(**CFG, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=3))
    want, _ = o.log_loss(synth_inputs('noise:8x8x3', B, seed=4).astype(np.float64))
    bounds = sorted(r[2] for r in res)
    assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == B
    for _, four, _ in res:
        np.testing.assert_allclose(four, want, rtol=1e-12)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


# ---- data-parallel training: one all-reduce of the flat gradient buffer -----------------------
def _flat(grads):
    return np.concatenate([np.asarray(g[net][k], np.float64).reshape(-1) for g in grads for net in ('A', 'b')
                           for k in sorted(g[net])])


def _grad_worker(rank, world, port, B, q):
    from oracle.grad_torch import loss_and_grads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    o = FlowOracle(**CFG, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=3))
    x = synth_inputs('noise:8x8x3', B, seed=4).astype(np.float64)
    lo, hi = shard_bounds(B, rank, world)
    _, grads = loss_and_grads(o, x[lo:hi])
    flat = torch.from_numpy(_flat(grads))
    allreduce_mean_gradients(flat, hi - lo)
    q.put((rank, flat.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 5])      # even and uneven shards
def test_two_rank_gradient_allreduce_equals_full_batch_gradient(B):
    from oracle.grad_torch import loss_and_grads
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = FlowOracle(**CFG, dtype=torch.float64)
    o.set_weights(init_weights(o.plan, 'rand', seed=3))
    _, want = loss_and_grads(o, synth_inputs('noise:8x8x3', B, seed=4).astype(np.float64))
    want = _flat(want)
    for _, got in res:
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    np.testing.assert_array_equal(res[0][1], res[1][1])      # every rank applies the identical update


# ---- replicas that start different are made identical before the first data-parallel update -----------------
def _sync_worker(rank, world, port, q):
    from arl_conditional_normalizing_flows_b200.conv_cINN_make_model import Adam, cFlow
    from arl_conditional_normalizing_flows_b200.sharding import replicas_checksum_equal, sync_replicas
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    np.random.seed(100 + rank)
    m = cFlow(**CFG, device="cpu")                  # default init: every rank draws different orthogonal kernels
    m.params.add_(float(rank))                       # ... and make sure of it
    m.compile(optimizer=Adam(3e-4))
    if rank == 0:                                    # rank 0 has optimizer state, rank 1 has none
        m.optimizer._m = torch.full_like(m.params, 0.25)
        m.optimizer._v = torch.full_like(m.params, 0.5)
        m.optimizer.iterations = 7
    before = replicas_checksum_equal(m)
    did = sync_replicas(m)
    after = replicas_checksum_equal(m)
    q.put((rank, before, did, after, m.params[:64].clone().numpy(), m.optimizer.iterations,
           float(m.optimizer._m[0]), float(m.optimizer._v[-1])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicas_are_synchronised_before_training():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, before, did, after, _, it, m0, v0 in res:
        assert before is False and did is True and after is True
        assert it == 7 and m0 == 0.25 and v0 == 0.5
    np.testing.assert_array_equal(res[0][4], res[1][4])


# ---- the overlapped, bucketed form of the gradient all-reduce (sharding.BucketedGradAllReduce) ----------------------
def _bucket_worker(rank, world, port, n_local, bucket_bytes, q):
    from arl_conditional_normalizing_flows_b200.sharding import BucketedGradAllReduce
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(10 + rank)
    local = torch.randn(1000, generator=g, dtype=torch.float32)
    flat = local.clone()
    allreduce_mean_gradients(flat, n_local[rank])
    # the backward pass reports the coupling layers from the last to the first; [0, 40) and [900, 1000) are never
    # reported (finish() must still reduce them), the layer slices are uneven
    layers = [(40, 150), (200, 100), (300, 350), (650, 250)]        # (offset, count); [190, 200) is a hole as well
    buf = local.clone()
    red = BucketedGradAllReduce(buf, n_local[rank], bucket_bytes=bucket_bytes)
    for li in reversed(range(len(layers))):
        red.layer_ready(li, *layers[li])
    red.finish()
    q.put((rank, flat.numpy(), buf.numpy(), list(red.buckets)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_local,bucket_bytes", [((3, 3), 1200), ((4, 1), 1), ((2, 5), 1 << 20)])
def test_two_rank_bucketed_allreduce_equals_flat_allreduce(n_local, bucket_bytes):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, n_local, bucket_bytes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, flat, bucketed, buckets in res:
        np.testing.assert_allclose(bucketed, flat, rtol=1e-6, atol=1e-7)
        # every element went through exactly one collective
        cover = np.zeros(1000, np.int32)
        for lo, hi in buckets:
            cover[lo:hi] += 1
        assert (cover == 1).all()
    np.testing.assert_array_equal(res[0][2], res[1][2])
    first = res[0][3]
    if bucket_bytes == 1:           # one bucket per layer, issued in backward order, then the unreported ranges
        assert first[:4] == [(650, 900), (300, 650), (200, 300), (40, 190)]
    if bucket_bytes == 1 << 20:     # everything reported merges into as few buckets as contiguity allows
        assert first[0] == (200, 900)


def test_bucketed_allreduce_single_process_is_a_no_op_and_reraises_hook_errors():
    from arl_conditional_normalizing_flows_b200.sharding import BucketedGradAllReduce
    g = torch.arange(10, dtype=torch.float32)
    red = BucketedGradAllReduce(g, 4)
    red.layer_ready(0, 0, 10)
    assert red.finish() is g and torch.equal(g, torch.arange(10, dtype=torch.float32))
    red = BucketedGradAllReduce(g, 4)
    red.active = True                       # force the bucket path without a process group: the error must surface
    red.layer_ready(0, 0, 10)
    with pytest.raises(Exception):
        red.finish()
