"""Double-buffered host -> device staging of a step's inputs (the reference does
``inputs[key] = ipt.to(self.device)`` synchronously at trainer.py:373-374).

``HostBatchStager`` owns ``depth`` device-resident copies of the batch dictionaries and a copy stream.
``submit(host_batch)`` enqueues the H2D copies of the NEXT batch on the copy stream (the host tensors must
be pinned for the copies to overlap), ``take()`` makes the compute stream wait for the oldest submitted
batch and returns its device tensors.  With one batch in flight the copies of step i+1 overlap the loss
kernels of step i, so a step costs max(copy, compute) instead of their sum.

``submit(host_batch, post=fn)`` also runs ``fn(device_batch)`` on the copy stream right behind the copies:
the on-GPU input pipeline (``input_pipeline.LossInputPipeline``: 8-bit frames -> colour pyramid) goes there,
so its small, latency-bound kernels run next to the previous step's loss kernels (the copy stream has the
higher priority, its blocks take the first SM slots that free up) instead of in front of this step's.
"""
from __future__ import annotations

import collections

import torch


class HostBatchStager:
    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.slots = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [None] * depth      # recorded on the compute stream when a slot's consumer is done
        self.queue = collections.deque()
        self.next_slot = 0
        self.bytes_per_batch = 0

    def _alloc_like(self, host_batch):
        return {k: torch.empty_like(v, device=self.device) for k, v in host_batch.items()}

    def submit(self, host_batch, post=None):
        """Enqueue the copies of one batch (dict of pinned CPU tensors), then ``post(device_batch)``."""
        if len(self.queue) >= self.depth:
            raise RuntimeError("all %d staging slots are in flight; call take() first" % self.depth)
        k = self.next_slot
        self.next_slot = (k + 1) % self.depth
        if self.slots[k] is None:
            self.slots[k] = self._alloc_like(host_batch)
            # the caching allocator hands out blocks in compute-stream order: the block may have just been freed
            # by a tensor whose kernels are still queued there, so the first copy into it waits for that stream
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        dst = self.slots[k]
        with torch.no_grad(), torch.cuda.stream(self.copy_stream):
            if self.free[k] is not None:
                self.copy_stream.wait_event(self.free[k])  # do not overwrite a slot still being read
            for key, v in host_batch.items():
                dst[key].copy_(v, non_blocking=True)
            if post is not None:
                post(dst)
            self.ready[k].record(self.copy_stream)
        self.bytes_per_batch = sum(v.numel() * v.element_size() for v in host_batch.values())
        self.queue.append(k)

    def take(self):
        """Device tensors of the oldest submitted batch; the current stream waits for its copies."""
        k = self.queue.popleft()
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        self._last = k
        return self.slots[k]

    def release(self):
        """Mark the batch returned by the last take() as consumed.  Call it only after EVERY kernel that reads
        the slot has been enqueued — including the backward of the loss step, which reads K and the gradient
        buffers through raw pointers (functional._FusedLoss) — or the next submit() may overwrite them."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[self._last] = ev
