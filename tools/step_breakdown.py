"""Where the step time outside k_photometric goes: CUDA-graph replays of growing prefixes of the C1 step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from unsupervised_pose_estimation_b200 import synthetic  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
cfg = dict(synthetic.CONFIGS["C1"])
dev = torch.device("cuda", 0)
wl = bench.Workload(cfg, "smooth", dev, 2)
st = wl.sets[0]
inputs, leaves = st["inputs"], st["leaves"]
path = wl.path
plan = path._vsl_plan(torch.float32)


def capture(fn, warm=3):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    return g, keep


def time_graph(g, n=300):
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best * 1e3


def noise_serial():
    return [torch.randn(12, 2, 192, 640, device=dev) for _ in range(4)]


def noise_parallel():
    return path._vsl_draw_noise(plan, dev, 4)


def fwd():
    out = dict(leaves)
    return path.compute_losses(inputs, out)


def full():
    out = dict(leaves)
    losses = path.compute_losses(inputs, out)
    return torch.autograd.grad(losses["loss"], list(leaves.values()))


def empty():
    return torch.empty(1, device=dev).zero_()


res = {}
for name, fn in (("one tiny kernel (graph launch floor)", empty), ("4 x randn, one stream", noise_serial),
                 ("4 x randn, parallel branches", noise_parallel), ("noise + k_photometric + k_epilogue", fwd),
                 ("full step (+ k_combine)", full)):
    g, keep = capture(fn)
    res[name] = time_graph(g)
    print("%-45s %8.1f us" % (name, res[name]), flush=True)

from unsupervised_pose_estimation_b200.graph import GraphedLossStep  # noqa: E402
for pf in (False, True):
    gs = GraphedLossStep(path, inputs, leaves, noise_prefetch=pf)
    for _ in range(20):
        gs.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(300):
            gs.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 300)
    print("%-45s %8.1f us" % ("GraphedLossStep(noise_prefetch=%s)" % pf, best * 1e3), flush=True)
