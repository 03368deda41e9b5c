"""Developer probe: cost of the reference-faithful side outputs (vsl_side_outputs="eager") at config 1."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from unsupervised_pose_estimation_b200 import synthetic

dev = torch.device("cuda", 0)
cfg = dict(synthetic.CONFIGS["C1"])
wl = bench.Workload(cfg, "smooth", dev, 2)
for mode in ("none", "eager"):
    wl.path.vsl_side_outputs = mode
    for i in range(5):
        wl.step(wl.sets[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        wl.step(wl.sets[i % 2])
    e1.record()
    torch.cuda.synchronize()
    print("side outputs %-5s: %.3f ms per eager step" % (mode, e0.elapsed_time(e1) / 50))
