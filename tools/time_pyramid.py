"""Developer probe: device time of the input-pipeline kernels at config 1 (run on the GPU box)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unsupervised_pose_estimation_b200.input_pipeline import FramePyramid

graphs = "--no-graph" not in sys.argv
for name, levels in (("target: 4 levels", None), ("source: level 0", [0])):
    pyr = FramePyramid(12, 192, 640, 4, "cuda", levels=levels)
    x = torch.randint(0, 256, (12, 192, 640, 3), dtype=torch.uint8, device="cuda")
    for _ in range(5):
        pyr(x)
    torch.cuda.synchronize()
    if not graphs:
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            pyr(x)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(name, "%.1f us per call" % (e0.elapsed_time(e1) / 20 * 1000))
