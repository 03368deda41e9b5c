"""Developer probe: the loss step under torch.no_grad() (forward-only kernel) against the differentiable step."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from unsupervised_pose_estimation_b200 import synthetic

dev = torch.device("cuda", 0)
wl = bench.Workload(dict(synthetic.CONFIGS["C1"]), "smooth", dev, 2)


def fwd(st):
    with torch.no_grad():
        outputs = dict(st["leaves"])
        wl.path.generate_images_pred(st["inputs"], outputs)
        return wl.path.compute_losses(st["inputs"], outputs)


for name, fn in (("fwd+bwd", lambda st: wl.step(st)), ("forward only (no_grad)", fwd)):
    for i in range(5):
        fn(wl.sets[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        fn(wl.sets[i % 2])
    e1.record()
    torch.cuda.synchronize()
    print("%-24s %.3f ms per eager step" % (name, e0.elapsed_time(e1) / 50))

from unsupervised_pose_estimation_b200 import functional as VF
for name, fn in (("fwd+bwd", lambda st: wl.step(st)), ("forward only (no_grad)", fwd)):
    ev = VF.KernelEvents()
    wl.path._vsl_plan().kernel_events = ev
    for i in range(20):
        fn(wl.sets[i % 2])
    torch.cuda.synchronize()
    ms = ev.drain_ms()
    wl.path._vsl_plan().kernel_events = None
    print("%-24s k_photometric %.3f ms" % (name, sum(ms) / len(ms)))
