"""CUDA-graph replay of the whole loss step (warp + losses + backward to the leaves).

The step is ~25 small host-side launches around one big kernel; enqueuing them from Python costs about as
much as the GPU work.  Shapes are static (the reference fixes batch/height/width in
``BackprojectDepth(batch,h,w)``), so the step is captured once into a ``torch.cuda.CUDAGraph`` and
replayed: one host call per step.  The tie-break ``torch.randn`` draws are captured too; PyTorch's
graph-safe generator advances the global RNG stream on every replay exactly as the eager calls would.

    step = GraphedLossStep(path, inputs, leaves)     # static device buffers, leaves require grad
    inputs[("color", 0, 0)].copy_(new_batch)          # write new data into the SAME tensors
    losses, grads = step.replay()                     # static output tensors, valid until the next replay
"""
from __future__ import annotations

import torch


class GraphedLossStep:
    """``noise_prefetch=True`` (opt-in, see ``ViewSynthesisLossMixin.vsl_noise_prefetch``): the step is captured twice
    over two static noise sets -- one graph consumes set A while it draws set B, the other the reverse -- and
    ``replay()`` alternates, so the generator kernels of step k+1 run in the shadow of step k's loss kernels.  The
    k-th replay uses the k-th group of draws of the global generator, exactly like the un-pipelined step; the first
    group is drawn in the constructor."""

    def __init__(self, path, inputs, leaves, loss_key="loss", warmup=3, pre=None, noise_prefetch=False,
                 capture_priority=-1):
        self.path, self.inputs, self.leaves = path, inputs, leaves
        self.keys = list(leaves.keys())
        dev = next(iter(leaves.values())).device
        # one plan, or one per level with --v1_multiscale; built for the dtype the images are stored in
        plans = path._vsl_level_plans(inputs[("color", 0, 0)].dtype)
        for plan, _, _ in plans:
            if plan.kernel_events is not None:
                raise RuntimeError("kernel timing events cannot be recorded inside a captured graph")
        single = path._vsl_plan(inputs[("color", 0, 0)].dtype)
        self.noise_prefetch = bool(noise_prefetch) and not isinstance(single, list) and bool(single.automask)
        self._noise = None
        if self.noise_prefetch:
            shape = (single.batch, single.noise_channels, single.height, single.width)
            self._noise = [[torch.empty(shape, dtype=torch.float32, device=dev) for _ in path.opt.scales] for _ in range(2)]
            for t in self._noise[0]:
                t.normal_()   # the first step's draws, in the reference's order

        def run(phase=0):
            if self.noise_prefetch:
                path.vsl_noise_prefetch = True
                path._vsl_noise_ahead, path._vsl_noise_out = self._noise[phase], self._noise[1 - phase]
            try:
                if pre is not None:
                    pre()   # e.g. the on-GPU input pipeline (input_pipeline.LossInputPipeline) filling `inputs`
                outputs = dict(leaves)
                path.generate_images_pred(inputs, outputs)
                losses = path.compute_losses(inputs, outputs)
                grads = torch.autograd.grad(losses[loss_key], [leaves[k] for k in self.keys], allow_unused=True)
            finally:
                if self.noise_prefetch:
                    path.vsl_noise_prefetch = False
                    path._vsl_noise_ahead = path._vsl_noise_out = None
            return outputs, losses, grads

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):   # warm-up off the default stream: builds the plan, calibrates, sizes the pools
            for k in range(warmup):
                run(k % 2)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._steps = []
        first = warmup % 2   # the set the last warm-up run drew into
        # with the pipelined noise the loss chain is captured on a HIGH-priority stream: the generator kernels (side
        # streams, default = lowest priority) become eligible together with k_photometric, and the CTA scheduler
        # must hand the SMs to the loss kernel first and fit the generator's CTAs into what its last wave leaves idle
        # (measured at C1: 0.773 ms per step with priority -1, 0.778 ms on a default-priority capture stream)
        cap_stream = (torch.cuda.Stream(device=dev, priority=capture_priority)
                      if self.noise_prefetch and capture_priority != 0 else None)
        for k in range(2 if self.noise_prefetch else 1):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap_stream):
                outputs, losses, grads = run((first + k) % 2)
            # static [2S+1] vector behind the loss dict (min_loss/s..., loss/s..., loss): one D2H copy reads it all
            self._steps.append((graph, outputs, losses, dict(zip(self.keys, grads)), getattr(path, "vsl_last_loss_vector", None)))
        self._next = 0
        self.graph, self.outputs, self.losses, self.grads, self.loss_vector = self._steps[0]

    def replay(self):
        """Run the step; ``outputs`` / ``losses`` / ``grads`` / ``loss_vector`` are static tensors, valid until the
        next replay (with noise_prefetch the two captured graphs own separate ones: use the attributes / return
        value of the call, not references kept from an earlier one)."""
        self.graph, self.outputs, self.losses, self.grads, self.loss_vector = self._steps[self._next]
        self._next = (self._next + 1) % len(self._steps)
        self.graph.replay()
        return self.losses, self.grads
