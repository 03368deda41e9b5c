"""Batch-sharded data parallelism for the loss path.

The path shards by image: every term is per-pixel or per-image followed by a mean over the batch, so
ranks need no exchange inside the path.  What crosses ranks is outside it: the loss scalars (for
logging) and the depth/pose-network gradients (one bucketed all-reduce).  One process per GPU,
``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests).  The reference has no distributed code
at all; each rank builds its modules with its LOCAL batch size because the
reference bakes ``batch_size`` into ``BackprojectDepth`` / ``Project3D`` (layers.py:225-232).
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


def all_reduce_grads(params, bucket_bytes=64 << 20, group=None):
    """Average ``.grad`` of the given parameters over ranks in flat buckets (the only data-path
    collective of config C5: 28.6 M depth/pose-net parameters, 114.6 MB fp32).  Parameters without a
    gradient (torchvision's unused ``fc``) are skipped on every rank alike."""
    if not dist.is_initialized():
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    n_buckets, i = 0, 0
    while i < len(grads):
        bucket, size = [], 0
        while i < len(grads) and (not bucket or size + grads[i].numel() * grads[i].element_size() <= bucket_bytes):
            bucket.append(grads[i])
            size += grads[i].numel() * grads[i].element_size()
            i += 1
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= world
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_buckets += 1
    return n_buckets
