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
