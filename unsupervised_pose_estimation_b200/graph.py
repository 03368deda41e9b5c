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
    def __init__(self, path, inputs, leaves, loss_key="loss", warmup=3, pre=None):
        self.path, self.inputs, self.leaves = path, inputs, leaves
        self.keys = list(leaves.keys())
        dev = next(iter(leaves.values())).device
        # one plan, or one per level with --v1_multiscale; built for the dtype the images are stored in
        for plan, _, _ in path._vsl_level_plans(inputs[("color", 0, 0)].dtype):
            if plan.kernel_events is not None:
                raise RuntimeError("kernel timing events cannot be recorded inside a captured graph")

        def run():
            if pre is not None:
                pre()   # e.g. the on-GPU input pipeline (input_pipeline.LossInputPipeline) filling `inputs`
            outputs = dict(leaves)
            path.generate_images_pred(inputs, outputs)
            losses = path.compute_losses(inputs, outputs)
            grads = torch.autograd.grad(losses[loss_key], [leaves[k] for k in self.keys], allow_unused=True)
            return outputs, losses, grads

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):   # warm-up off the default stream: builds the plan, calibrates, sizes the pools
            for _ in range(warmup):
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs, self.losses, grads = run()
        self.grads = dict(zip(self.keys, grads))
        # static [2S+1] vector behind the loss dict (min_loss/s..., loss/s..., loss): one D2H copy reads it all
        self.loss_vector = getattr(path, "vsl_last_loss_vector", None)

    def replay(self):
        self.graph.replay()
        return self.losses, self.grads
