"""Batch-sharded data parallelism for the loss path.

The path shards by image: every term is per-pixel or per-image followed by a mean over the batch, so
ranks need no exchange inside the path.  What crosses ranks is outside it: the loss scalars (for
logging) and the depth/pose-network gradients.  One process per GPU, ``torch.distributed`` (NCCL on
GPUs; gloo in the CPU tests).  The reference has no distributed code at all (its step is
trainer.py:297-343: forward, ``zero_grad``, ``backward``, ``step``); each rank builds its modules with its
LOCAL batch size because the reference bakes ``batch_size`` into ``BackprojectDepth`` / ``Project3D``
(layers.py:225-232).

``GradBuckets`` is the gradient exchange: the parameters' ``.grad`` are views into one persistent flat
buffer, cut into buckets in reverse parameter order (the order backward produces them); a
post-accumulate-grad hook per parameter launches a bucket's all-reduce on a side stream the moment its last
gradient has been written, so the exchange overlaps the rest of the backward instead of following it, with no
``cat`` before and no copy-back after.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Contiguous slice [lo, hi) of the global batch owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world):
    """Slice every tensor of a dict (inputs / outputs / leaves) along dim 0 to this rank's shard."""
    out = {}
    for k, v in tensors.items():
        lo, hi = shard_range(v.shape[0], rank, world)
        out[k] = v[lo:hi]
    return out


def all_reduce_losses(losses, local_batch, group=None):
    """Global-batch means from per-rank means: sum_r (B_r / B) * loss_r.  One small all-reduce."""
    if not dist.is_initialized():
        return losses
    keys = sorted(losses)
    vec = torch.stack([losses[k].detach() for k in keys]).double() * float(local_batch)
    tot = torch.tensor([float(local_batch)], dtype=torch.float64, device=vec.device)
    buf = torch.cat([vec, tot])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return {k: (buf[i] / buf[-1]).float() for i, k in enumerate(keys)}


class GradBuckets:
    """Overlapped, bucketed all-reduce of network gradients into the GLOBAL-batch mean.

        buckets = GradBuckets(params, local_batch=B_r, global_batch=B)   # once, after the nets are on the device
        for batch in loader:
            buckets.begin_step()             # zeroes the flat buffer, (re-)attaches the .grad views
            loss.backward()                  # hooks launch each bucket's all-reduce as soon as it is complete
            buckets.finish()                 # buckets whose hooks never fired (unused parameters), then wait
            optimizer.step()                 # do NOT call optimizer.zero_grad(set_to_none=True) instead of begin_step()

    * The bucket layout is a function of the fixed parameter list only, so it is identical on every rank even if
      some rank produces no gradient for a parameter (torchvision's unused ``fc``): such a parameter contributes
      the zeros ``begin_step`` wrote.
    * Each rank's gradients are d(local mean)/dw; they are weighted by ``local_batch / global_batch`` before the
      SUM, which is the gradient of the global-batch mean also when shards differ in size by one
      (``shard_range``) — the same weighting ``all_reduce_losses`` applies.
    * On CUDA the all-reduce of a bucket is enqueued on ``comm_stream`` behind an event recorded where the hook
      fires; ``finish`` makes the current stream wait for all of them.  On CPU (gloo, tests) the same calls run
      with ``async_op=True`` and are waited for in ``finish``.
    """

    def __init__(self, params, local_batch=1, global_batch=None, bucket_bytes=32 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters require grad")
        p0 = self.params[0]
        if any(p.device != p0.device or p.dtype != p0.dtype for p in self.params):
            raise ValueError("GradBuckets needs all parameters on one device with one dtype")
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        gb = float(global_batch if global_batch is not None else local_batch * self.world)
        self.weight = float(local_batch) / gb
        self.is_cuda = p0.is_cuda
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=p0.dtype, device=p0.device)
        # offsets in REVERSE parameter order: backward reaches the last layers first, so bucket 0 fills first
        self.offset, off = {}, 0
        for p in reversed(self.params):
            self.offset[id(p)] = off
            off += p.numel()
        self.bucket_of, self.bounds = {}, []
        lo, cur = 0, 0
        esz = p0.element_size()
        for p in reversed(self.params):
            n = p.numel()
            if cur > lo and (cur + n - lo) * esz > bucket_bytes:
                self.bounds.append((lo, cur))
                lo = cur
            self.bucket_of[id(p)] = len(self.bounds)
            cur += n
        self.bounds.append((lo, cur))
        self.members = [0] * len(self.bounds)
        for p in self.params:
            self.members[self.bucket_of[id(p)]] += 1
        self.comm_stream = torch.cuda.Stream(device=p0.device) if self.is_cuda else None
        self.pending, self.launched, self.work = [], [], []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.begin_step()

    @property
    def num_buckets(self):
        return len(self.bounds)

    def grad_view(self, p):
        off = self.offset[id(p)]
        return self.flat[off:off + p.numel()].view_as(p)

    def begin_step(self):
        """Zero the flat buffer and make every ``p.grad`` the view into it (``zero_grad(set_to_none=True)`` or a
        fresh ``torch.autograd.grad`` would have dropped the views; AccumulateGrad adds in place into a defined
        grad, so the views survive the backward)."""
        if self.work:
            self.finish()
        self.flat.zero_()
        for p in self.params:
            v = self.grad_view(p)
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v
        self.pending = list(self.members)
        self.launched = [False] * len(self.bounds)

    def _launch(self, b):
        lo, hi = self.bounds[b]
        bucket = self.flat[lo:hi]
        self.launched[b] = True
        if self.world == 1:
            return
        if self.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(bucket.device))
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                if self.weight != 1.0:
                    bucket.mul_(self.weight)
                self.work.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            if self.weight != 1.0:
                bucket.mul_(self.weight)
            self.work.append(dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _on_grad(self, p):
        b = self.bucket_of[id(p)]
        if p.grad is not None and p.grad.data_ptr() != self.grad_view(p).data_ptr():
            # someone replaced the view (zero_grad(set_to_none=True) before backward): fold the value back in
            self.grad_view(p).add_(p.grad)
            p.grad = self.grad_view(p)
        self.pending[b] -= 1
        if self.pending[b] == 0 and not self.launched[b]:
            self._launch(b)

    def finish(self):
        """Launch the buckets that are still incomplete (parameters without a gradient this step hold zeros),
        then wait: afterwards every ``p.grad`` is the global-batch-mean gradient.  Returns the bucket count."""
        for b in range(len(self.bounds)):
            if not self.launched[b]:
                self._launch(b)
        for w in self.work:
            w.wait()   # CUDA: the current stream waits for the collective; CPU: blocks
        self.work = []
        return len(self.bounds)

    def remove_hooks(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def all_reduce_grads(params, bucket_bytes=64 << 20, group=None, local_batch=1, global_batch=None):
    """Blocking form (after backward): global-batch-mean ``.grad`` over ranks in flat buckets.  The bucket
    layout covers EVERY parameter that requires grad — a missing gradient travels as zeros — so ranks can never
    disagree on sizes.  Prefer ``GradBuckets``, which overlaps the exchange with the backward."""
    if not dist.is_initialized():
        return 0
    params = [p for p in params if p.requires_grad]
    world = dist.get_world_size(group)
    weight = float(local_batch) / float(global_batch if global_batch is not None else local_batch * world)
    n_buckets, i = 0, 0
    while i < len(params):
        bucket, size = [], 0
        while i < len(params) and (not bucket or size + params[i].numel() * params[i].element_size() <= bucket_bytes):
            bucket.append(params[i])
            size += params[i].numel() * params[i].element_size()
            i += 1
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        flat *= weight
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        n_buckets += 1
    return n_buckets
