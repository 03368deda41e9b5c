"""Developer probe: how much do concurrent H2D copies slow the graph-replayed loss step?"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from unsupervised_pose_estimation_b200 import synthetic
from unsupervised_pose_estimation_b200.graph import GraphedLossStep

dev = torch.device("cuda", 0)
cfg = dict(synthetic.CONFIGS["C1"])
wl = bench.Workload(cfg, "smooth", dev, 4)
graphs = [GraphedLossStep(wl.path, st["inputs"], st["leaves"]) for st in wl.sets]
side = torch.cuda.Stream(device=dev, priority=-1)
ev = lambda: torch.cuda.Event(enable_timing=True)
for mb in (0, 5, 10, 21, 42):
    n = mb * 1000 * 1000
    host = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()
    dst = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for i in range(100):
            graphs[i % 4].replay()
            if n:
                with torch.cuda.stream(side):
                    dst.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
    print("H2D %2d MB per step: %.4f ms per step" % (mb, e0.elapsed_time(e1) / 100))

# the same with the e2e loop's cross-stream dependencies: step i waits for the copy issued during step i-1,
# the copy waits for step i-2 (slot re-use)
n = 21 * 1000 * 1000
host = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
for deps in ("none", "step waits copy", "step waits copy, copy waits step"):
    for rep in range(2):
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [None, None]
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        with torch.cuda.stream(side):
            dst[0].copy_(host, non_blocking=True)
            ready[0].record(side)
        for i in range(100):
            k = i % 2
            if deps != "none":
                torch.cuda.current_stream().wait_event(ready[k])
            graphs[i % 4].replay()
            f = torch.cuda.Event(); f.record(); free[k] = f
            with torch.cuda.stream(side):
                if deps.endswith("copy waits step") and free[1 - k] is not None:
                    side.wait_event(free[1 - k])
                dst[1 - k].copy_(host, non_blocking=True)
                ready[1 - k].record(side)
        e1.record()
        torch.cuda.synchronize()
    print("deps = %-36s %.4f ms per step" % (deps, e0.elapsed_time(e1) / 100))
