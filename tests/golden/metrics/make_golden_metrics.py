"""Golden values of the reference's depth metrics / SLlog, from the UNMODIFIED reference (build container only):

    python tests/golden/metrics/make_golden_metrics.py   ->  tests/golden/metrics/metrics.npz

`SLlog` and `compute_depth_errors` come from /root/reference/layers.py as they are; `compute_depth_losses` is
`Trainer.compute_depth_losses` (trainer.py:688-716) bound on a namespace with the one attribute it reads
(`depth_metric_names`).  Inputs are re-generated from seeds by oracle.metrics_oracle.metric_inputs, so the
fixture holds outputs only.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import load_reference  # noqa: E402
from oracle import metrics_oracle as M  # noqa: E402


def main():
    torch.set_num_threads(1)
    ref_layers, ref_trainer = load_reference()
    blob = {}
    for seed in (0, 1):
        fake, real = M.metric_inputs(seed, "sllog")
        fake = fake.clone().requires_grad_(True)
        real = real.clone().requires_grad_(True)
        loss = ref_layers.SLlog()(fake, real)
        gf, gr = torch.autograd.grad(loss, [fake, real])
        blob["sllog|%d|loss" % seed] = loss.detach().numpy()
        blob["sllog|%d|grad_fake" % seed] = gf.numpy()
        blob["sllog|%d|grad_real" % seed] = gr.numpy()
        gt, pred = M.metric_inputs(seed, "errors")
        blob["errors|%d" % seed] = np.array([float(v) for v in ref_layers.compute_depth_errors(gt, pred)], np.float64)
        dpred, dgt = M.metric_inputs(seed, "depth_losses")
        ns = types.SimpleNamespace(depth_metric_names=[
            "de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"])   # trainer.py:255-256
        losses = {}
        types.MethodType(ref_trainer.Trainer.compute_depth_losses, ns)({"depth_gt": dgt}, {("depth", 0, 0): dpred}, losses)
        blob["depth_losses|%d" % seed] = np.array([float(losses[k]) for k in ns.depth_metric_names], np.float64)
    blob["meta|torch"] = np.array(torch.__version__)
    path = os.path.join(HERE, "metrics.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in sorted(blob):
        if blob[k].size <= 8:
            print(k, blob[k])


if __name__ == "__main__":
    main()
