"""One process per library build: quick parity check against the oracle, then step / kernel timing at a config.

    VSL_LIB_PATH=variants/libvsl_x.so python tools/variant_probe.py [--config C1] [--steps 200] [--bf16] [--no-parity]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from unsupervised_pose_estimation_b200 import functional as VF, layers as L, synthetic  # noqa: E402
from unsupervised_pose_estimation_b200.graph import GraphedLossStep  # noqa: E402
from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt  # noqa: E402


def parity(cfg, bf16):
    from oracle import vsl_oracle as O
    B, H, W, frames = 2, cfg["height"], cfg["width"], cfg["frame_ids"]
    worst = {"loss": 0.0, "grad": 0.0, "mask_flips": 0}
    for family, seed in (("smooth", 3), ("iid", 4)):
        opt = O.make_opt(height=H, width=W, batch_size=B, frame_ids=list(frames))
        inputs, outputs, leaves = synthetic.make_batch(B, H, W, frames, cfg["K"], seed=seed, family=family, device="cuda")
        if bf16:
            inputs = {k: (v.bfloat16() if k[0] == "color" else v) for k, v in inputs.items()}
        in32 = {k: (v.float() if k[0] == "color" else v) for k, v in inputs.items()}

        def poses(fn):
            out = dict(outputs)
            for f in frames[1:]:
                if f != "s":
                    out[("cam_T_cam", 0, f)] = fn(leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
            return out
        path = LossPath(make_opt(**vars(opt)), device="cuda", side_outputs="none")
        out = poses(L.transformation_from_parameters)
        torch.manual_seed(1)
        losses = path.compute_losses(inputs, out)
        g = torch.autograd.grad(losses["loss"], list(leaves.values()))
        ref_out = poses(O.transformation_from_parameters)
        torch.manual_seed(1)
        ref = O.loss_step(opt, in32, ref_out)
        rg = torch.autograd.grad(ref["loss"], list(leaves.values()))
        for k in ref:
            worst["loss"] = max(worst["loss"], abs(losses[k].item() - ref[k].item()) / abs(ref[k].item()))
        for s in opt.scales:
            worst["mask_flips"] += int((out["identity_selection/%d" % s] != ref_out["identity_selection/%d" % s]).sum())
        for a, b in zip(g, rg):
            worst["grad"] = max(worst["grad"], ((a - b).norm() / b.norm()).item())
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C1")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--family", default="smooth")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(synthetic.CONFIGS[args.config])
    res = {"lib": os.path.basename(os.environ.get("VSL_LIB_PATH", "libvsl_b200.so")), "config": args.config, "bf16": args.bf16}
    if not args.no_parity:
        res["parity"] = parity(cfg, args.bf16)
    dev = torch.device("cuda", 0)
    ring = 4
    wl = bench.Workload(cfg, args.family, dev, ring, bf16_images=args.bf16)
    wl.path.vsl_parallel_noise = not os.environ.get("VSL_SERIAL_NOISE")
    res["parallel_noise"] = wl.path.vsl_parallel_noise
    for i in range(5):
        wl.step(wl.sets[i % ring])
    torch.cuda.synchronize()
    ev = VF.KernelEvents()
    wl.path._vsl_plan().kernel_events = ev
    for i in range(40):
        wl.step(wl.sets[i % ring])
    torch.cuda.synchronize()
    k = ev.drain_ms()
    wl.path._vsl_plan().kernel_events = None
    res["kernel_ms"] = round(sum(k) / len(k), 4)
    graphs = [GraphedLossStep(wl.path, st["inputs"], st["leaves"]) for st in wl.sets]
    for i in range(20):
        graphs[i % ring].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(3):
        e0.record()
        for i in range(args.steps):
            graphs[i % ring].replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / args.steps)
    res["step_ms"] = round(best, 4)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
