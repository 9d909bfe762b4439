"""Batch sharding across GPUs: one process per GPU, weights replicated, the batch split
into contiguous slices.  Likelihood evaluation and sampling are per-sample (LayerNorm is per-sample,
F:350-360), so the data path has NO collective; the only cross-sample quantities are the batch means
of cFlow.log_loss (M:1325, M:1840), recovered exactly from per-shard sums with one tiny all-reduce.
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
