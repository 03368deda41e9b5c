"""GPU probe: which rounding order does eager PyTorch-CUDA use for each stage of the path?

Runs the oracle (test infrastructure) on cuda and compares, bit for bit, with the library under
every VSL_ARITH_* variant.  Writes gpurun_out/probe.json.  Developer tool, not part of the product.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import vsl_oracle as O  # noqa: E402
from unsupervised_pose_estimation_b200 import _lib, synthetic  # noqa: E402
from unsupervised_pose_estimation_b200 import functional as VF  # noqa: E402
from unsupervised_pose_estimation_b200 import layers as L  # noqa: E402
from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt  # noqa: E402

dev = "cuda"
out = {}


def mism(a, b):
    return int((a != b).sum().item()), a.numel()


def run(B, H, W, frame_ids, family, seed):
    opt = O.make_opt(height=H, width=W, batch_size=B, frame_ids=list(frame_ids))
    inputs, outputs, leaves = synthetic.make_batch(B, H, W, frame_ids, seed=seed, family=family, device=dev,
                                                   pose_fn=L.transformation_from_parameters)
    def fresh_outputs():
        o = dict(outputs)
        for f in frame_ids[1:]:
            if f != "s":
                o[("cam_T_cam", 0, f)] = L.transformation_from_parameters(
                    leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
        return o

    ref_out = fresh_outputs()
    O.generate_images_pred(opt, inputs, ref_out)
    torch.manual_seed(123)
    ref_losses = O.compute_losses(opt, inputs, ref_out)
    ref_losses["loss"].backward()
    ref_grads = {k: v.grad.clone() for k, v in leaves.items()}
    for v in leaves.values():
        v.grad = None
    res = {}
    variants = {"cuda_order": 0, "true_div": 1, "dot_nofma": 2, "dot_reverse": 4, "ups_right": 8, "ups_nofma": 16,
                "tap_nofma": 32, "mean_div": 64}
    for name, arith in variants.items():
        lp = LossPath(make_opt(**vars(opt)), side_outputs="eager", arith=arith)
        o2 = fresh_outputs()
        lp.generate_images_pred(inputs, o2)
        r = {}
        for s in opt.scales:
            r["depth/%d" % s] = mism(o2[("depth", 0, s)], ref_out[("depth", 0, s)])
            for f in frame_ids[1:]:
                r["sample/%s/%d" % (f, s)] = mism(o2[("sample", f, s)], ref_out[("sample", f, s)])
                r["color/%s/%d" % (f, s)] = mism(o2[("color", f, s)], ref_out[("color", f, s)])
        tgt = inputs[("color", 0, 0)]
        pred = ref_out[("color", frame_ids[1], 0)].detach()
        r["reproj"] = mism(VF.reprojection_loss(pred, tgt, arith=arith), O.reprojection_loss(pred, tgt))
        r["ssim"] = mism(VF.ssim(pred, tgt), O.ssim(pred, tgt))
        torch.manual_seed(123)
        losses = lp.compute_losses(inputs, o2)
        losses["loss"].backward()
        for s in opt.scales:
            k = "identity_selection/%d" % s
            r["mask/%d" % s] = mism(o2[k], ref_out[k])
        r["loss_rel"] = {k: abs(losses[k].item() - ref_losses[k].item()) / abs(ref_losses[k].item()) for k in losses}
        g = {}
        for k, v in leaves.items():
            gr = ref_grads[k]
            g[repr(k)] = [((v.grad - gr).norm() / gr.norm()).item(), ((v.grad - gr).abs().max() / gr.abs().max()).item()]
            v.grad = None
        r["grad_relL2_relMax"] = g
        res[name] = r
    return res


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = False
    print(torch.__version__, torch.cuda.get_device_name(0))
    _lib.load()
    out["small_iid"] = run(2, 64, 96, [0, -1, 1], "iid", 0)
    out["c1_smooth_b4"] = run(4, 192, 640, [0, -1, 1], "smooth", 1)
    out["c1_iid_b4"] = run(4, 192, 640, [0, -1, 1], "iid", 2)
    out["stereo_iid_b2"] = run(2, 192, 640, [0, -1, 1, "s"], "iid", 3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    for case, res in out.items():
        for name, r in res.items():
            bad = {k: v for k, v in r.items() if isinstance(v, tuple) and v[0] != 0}
            print(case, name, "mismatches:", bad)
            if name == "cuda_order":
                print("   loss_rel", r["loss_rel"])
                print("   grads", r["grad_relL2_relMax"])
