"""Developer probe: do the input-pipeline kernels (high-priority side stream) overlap the loss graph?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from unsupervised_pose_estimation_b200 import synthetic
from unsupervised_pose_estimation_b200.graph import GraphedLossStep
from unsupervised_pose_estimation_b200.input_pipeline import LossInputPipeline

dev = torch.device("cuda", 0)
cfg = dict(synthetic.CONFIGS["C1"])
wl = bench.Workload(cfg, "smooth", dev, 1)
st = wl.sets[0]
g = GraphedLossStep(wl.path, st["inputs"], st["leaves"])
pipe = LossInputPipeline(wl.opt, dev)
frames = {f: torch.randint(0, 256, (12, 192, 640, 3), dtype=torch.uint8, device=dev) for f in cfg["frame_ids"]}
side = torch.cuda.Stream(device=dev, priority=-1)
ev = lambda: torch.cuda.Event(enable_timing=True)
for mode in ("graph alone", "pyramid alone", "both"):
    res = []
    for _ in range(5):
        torch.cuda.synchronize()
        a0, a1, b0, b1 = ev(), ev(), ev(), ev()
        a0.record()
        if mode != "pyramid alone":
            g.replay()
        a1.record()
        with torch.cuda.stream(side):
            side.wait_event(a0)
            b0.record()
            if mode != "graph alone":
                pipe(frames)
            b1.record()
        torch.cuda.synchronize()
        res.append((a0.elapsed_time(a1), a0.elapsed_time(b0), a0.elapsed_time(b1)))
    print(mode, "graph %.3f ms | side start %.3f end %.3f ms" % res[-1])
