"""Batch sharding across GPUs: one process per GPU, weights replicated, the batch split
into contiguous slices.  Likelihood evaluation and sampling are per-sample (LayerNorm is per-sample,
F:350-360), so the data path has NO collective; the only cross-sample quantities are the batch means
of cFlow.log_loss (M:1325, M:1840), recovered exactly from per-shard sums with one tiny all-reduce.
Training (cFlow.train_step, M:1850-1880) is data-parallel: every rank differentiates the mean loss of its own
shard and ONE sum all-reduce of the flat gradient buffer (NCCL over NVLink on the GPUs, gloo in the CPU tests)
gives the gradient of the global-batch loss; every rank then applies the identical Adam update.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) slice of `n` samples owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_loss(ll_z, ll_y, logdet, group=None):
    """The reference 4-tuple (loss, z_loss, y_loss, detJ_loss) of the GLOBAL batch from this rank's
    per-sample vectors (cFlow.last_per_sample).  Sums are all-reduced in fp64 and divided by the global
    count, so uneven shards are weighted correctly and the result equals a single-device batch (Q1)."""
    acc = torch.stack([ll_z.double().sum(), ll_y.double().sum(), logdet.double().sum(),
                       torch.tensor(float(ll_z.numel()), dtype=torch.float64, device=ll_z.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    mz, my, md = (acc[:3] / acc[3]).tolist()
    return -(mz + my + md), -mz, -my, -md


def log_loss_sharded(model, xy_local, group=None):
    """cFlow.log_loss on this rank's shard, reduced to the global-batch scalars."""
    model.log_loss(xy_local)
    ps = model.last_per_sample
    return global_loss(ps['ll_z'], ps['ll_y'], ps['logdet'], group)


def allreduce_mean_gradients(grads, n_local, group=None):
    """In place: this rank's gradient of ITS shard's mean loss -> the gradient of the GLOBAL batch's mean loss,
    sum_r (n_r / N) g_r.  One sum all-reduce of the flat buffer plus an 8-byte one for N (shards may be uneven);
    no host synchronisation.  A single process (or no process group) leaves grads untouched."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    n = torch.tensor([float(n_local)], dtype=torch.float64, device=grads.device)
    total = n.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    grads.mul_((n / total).to(grads.dtype))
    dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)
    return grads
