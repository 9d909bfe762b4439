"""Batch sharding across GPUs: one process per GPU, weights replicated, the batch split
into contiguous slices.  Likelihood evaluation and sampling are per-sample (LayerNorm is per-sample,
F:350-360), so the data path has NO collective; the only cross-sample quantities are the batch means
of cFlow.log_loss (M:1325, M:1840), recovered exactly from per-shard sums with one tiny all-reduce.
Training (cFlow.train_step, M:1850-1880) is data-parallel: every rank differentiates the mean loss of its own
shard and ONE sum all-reduce of the flat gradient buffer (NCCL over NVLink on the GPUs, gloo in the CPU tests)
gives the gradient of the global-batch loss; every rank then applies the identical Adam update.  With
`BucketedGradAllReduce` that all-reduce is issued per bucket of coupling layers, in the order the backward pass completes
them, so the transfer of the upper layers' gradients overlaps the differentiation of the lower ones.
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


def sync_replicas(model, src=0, group=None):
    """Make every rank's replica identical to rank `src`'s: the flat parameter buffer and, when the optimizer has
    state, Adam's m / v / step count.  Data-parallel training only reproduces single-device training if the replicas
    START identical (cFlow.__init__ draws its Orthogonal(0.1) kernels from an unseeded generator, so they do not), so
    cFlow.train_step calls this once before its first data-parallel update.  No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return False
    dist.broadcast(model.params, src=src, group=group)
    opt = getattr(model, 'optimizer', None)
    if opt is not None:
        it = torch.tensor([int(getattr(opt, 'iterations', 0))], dtype=torch.int64, device=model.params.device)
        dist.broadcast(it, src=src, group=group)
        opt.iterations = int(it.item())
        has = torch.tensor([0 if getattr(opt, '_m', None) is None else 1], dtype=torch.int64, device=model.params.device)
        dist.broadcast(has, src=src, group=group)
        if int(has.item()):
            if opt._m is None or opt._m.shape != model.params.shape:
                opt._m = torch.zeros_like(model.params)
                opt._v = torch.zeros_like(model.params)
            dist.broadcast(opt._m, src=src, group=group)
            dist.broadcast(opt._v, src=src, group=group)
    return True


def replicas_checksum_equal(model, group=None):
    """True iff every rank holds bit-identical parameters (max and min of a 64-bit fold of the buffer agree)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return True
    h = model.params.detach().contiguous().view(torch.int32).to(torch.int64)
    w = torch.arange(1, h.numel() + 1, dtype=torch.int64, device=h.device)
    c = ((h * w) % 2147483629).sum().reshape(1)
    lo, hi = c.clone(), c.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool((lo == hi).item())


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


class BucketedGradAllReduce:
    """The gradient all-reduce of a data-parallel train_step, overlapped with the backward pass.

    libcnf's backward pass walks the coupling layers from the last to the first and reports, through the
    `cnf_layer_grads_ready_fn` hook of `cnf_flow_loss_and_grad_hooked`, every layer whose slice of the flat gradient
    buffer is complete on the compute stream.  `layer_ready` merges consecutive slices into buckets of at least
    `bucket_bytes` (sized for launch latency, not link count: NVSwitch gives every pair full bandwidth) and issues one
    asynchronous sum all-reduce per bucket; NCCL orders it behind the compute stream's position at that moment and runs it
    on its own stream while the layers below are still being differentiated.  `finish` reduces whatever the hook did not
    cover and makes the compute stream wait for every bucket.  The result equals `allreduce_mean_gradients` (the slices
    are scaled by n_local / N before the sum, so uneven shards are weighted correctly); a single process does nothing.
    """

    def __init__(self, grads, n_local, bucket_bytes=8 << 20, group=None):
        self.grads = grads
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.bucket_elems = max(1, int(bucket_bytes) // grads.element_size())
        self.works = []
        self.buckets = []            # (lo, hi) element ranges already handed to the collective, in issue order
        self._lo = self._hi = None   # pending contiguous range
        self.error = None            # an exception raised inside the ctypes callback (ctypes would only print it)
        if self.active:
            n = torch.tensor([float(n_local)], dtype=torch.float64, device=grads.device)
            total = n.clone()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)      # 8 bytes, queued before the backward pass
            self.scale = (n / total).to(grads.dtype)

    def layer_ready(self, layer, offset, count):
        """grads[offset, offset + count) of coupling layer `layer` is complete on the current stream."""
        if not self.active or self.error is not None:
            return
        try:
            lo, hi = int(offset), int(offset) + int(count)
            if self._lo is None:
                self._lo, self._hi = lo, hi
            elif hi == self._lo:            # the backward pass moves down the buffer
                self._lo = lo
            elif lo == self._hi:
                self._hi = hi
            else:
                self._flush()
                self._lo, self._hi = lo, hi
            if self._hi - self._lo >= self.bucket_elems:
                self._flush()
        except Exception as e:      # noqa: BLE001 -- re-raised by finish()
            self.error = e

    def _reduce(self, lo, hi):
        if hi <= lo:
            return
        sl = self.grads[lo:hi]
        sl.mul_(self.scale)
        self.works.append(dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.buckets.append((lo, hi))

    def _flush(self):
        if self._lo is not None:
            self._reduce(self._lo, self._hi)
            self._lo = self._hi = None

    def finish(self):
        """Reduce what the hook did not report (nothing, for a cFlow plan) and wait for every bucket."""
        if self.error is not None:
            raise self.error
        if not self.active:
            return self.grads
        self._flush()
        pos = 0
        for lo, hi in sorted(self.buckets):
            self._reduce(pos, lo)
            pos = max(pos, hi)
        self._reduce(pos, self.grads.numel())
        for w in self.works:
            w.wait()          # NCCL: the current stream waits for the collective's stream; gloo: blocks the host
        self.works = []
        return self.grads
