"""Where does the gradient error live?  (VERDICT r1 weak #1.)

For every BASELINE single-GPU config at its full batch: gradient of losses["loss"] with respect to every leaf
the path differentiates (disp_0..3, each cam_T_cam), from the fused CUDA path, the fp32 oracle and the fp64
oracle on the same GPU with the same tie-break noise; rel-L2 and max-norm errors per leaf, written as a
markdown table to gpurun_out/r2_grad_error_by_leaf.md.

    python tools/grad_error_by_leaf.py [C1 C2 ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vsl_oracle as O                                       # noqa: E402
from unsupervised_pose_estimation_b200 import layers as L, synthetic    # noqa: E402
from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt  # noqa: E402

DEV = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def leaves_for(cfg, seed, family, dtype=torch.float32):
    inputs, outputs, raw = synthetic.make_batch(cfg["batch"], cfg["height"], cfg["width"], cfg["frame_ids"], cfg["K"],
                                                seed=seed, family=family, device=DEV, requires_grad=False)
    leaves = {}
    for s in range(4):
        leaves[("disp", s)] = raw[("disp", s)].to(dtype).requires_grad_(True)
    for f in cfg["frame_ids"][1:]:
        if f != "s":
            T = L.transformation_from_parameters(raw[("axisangle", 0, f)][:, 0], raw[("translation", 0, f)][:, 0], f < 0)
            leaves[("cam_T_cam", 0, f)] = T.to(dtype).requires_grad_(True)
    inputs = {k: v.to(dtype) for k, v in inputs.items()}
    return inputs, leaves


def err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm()).item(), ((a - b).abs().max() / b.abs().max()).item()


def main():
    names = sys.argv[1:] or ["C1", "C2", "C3", "C4"]
    rows = []
    for name in names:
        for family in ("smooth", "iid"):
            cfg = synthetic.CONFIGS[name]
            opt = O.make_opt(height=cfg["height"], width=cfg["width"], batch_size=cfg["batch"],
                             frame_ids=list(cfg["frame_ids"]))
            F = len(cfg["frame_ids"]) - 1
            torch.manual_seed(5)
            noise = [torch.randn(cfg["batch"], F, cfg["height"], cfg["width"], device=DEV) for _ in opt.scales]
            grads = {}
            for tag, dtype in (("o32", torch.float32), ("o64", torch.float64)):
                inputs, leaves = leaves_for(cfg, 17, family, dtype)
                out = dict(leaves)
                O.generate_images_pred(opt, inputs, out)
                losses = O.compute_losses(opt, inputs, out, [z.to(dtype) for z in noise])
                grads[tag] = dict(zip(leaves, torch.autograd.grad(losses["loss"], list(leaves.values()))))
                masks = {s: out["identity_selection/%d" % s] for s in opt.scales}
                grads[tag + "_masks"] = masks
                del out, losses
                torch.cuda.empty_cache()
            inputs, leaves = leaves_for(cfg, 17, family)
            path = LossPath(make_opt(**vars(opt)), device=DEV, side_outputs="none")
            out = dict(leaves)
            torch.manual_seed(5)   # compute_losses draws the same noise tensors
            losses = path.compute_losses(inputs, out)
            grads["ours"] = dict(zip(leaves, torch.autograd.grad(losses["loss"], list(leaves.values()))))
            flips32 = sum(int((out["identity_selection/%d" % s] != grads["o32_masks"][s]).sum()) for s in opt.scales)
            flips64 = sum(int((grads["o32_masks"][s] != grads["o64_masks"][s].float()).sum()) for s in opt.scales)
            for k in leaves:
                e_ours32 = err(grads["ours"][k], grads["o32"][k])
                e_ours64 = err(grads["ours"][k], grads["o64"][k])
                e_o32_64 = err(grads["o32"][k], grads["o64"][k])
                rows.append((name, family, str(k), e_ours32, e_ours64, e_o32_64, flips32, flips64))
                print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_grad_error_by_leaf.md"), "w") as f:
        f.write("# Gradient error by leaf (full batch 12, same tie-break noise in all three runs)\n\n")
        f.write("rel-L2 = |a-b|_2/|b|_2, max = |a-b|_inf/|b|_inf.  `flips`: auto-mask pixels that differ (ours vs fp32 "
                "oracle; fp32 oracle vs fp64 oracle), summed over the four scales.\n\n")
        f.write("| config | images | leaf | ours vs o32 rel-L2 | max | ours vs o64 rel-L2 | max | o32 vs o64 rel-L2 | max | flips ours/o32 | flips o32/o64 |\n")
        f.write("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            f.write("| %s | %s | `%s` | %.2e | %.2e | %.2e | %.2e | %.2e | %.2e | %d | %d |\n"
                    % (r[0], r[1], r[2], r[3][0], r[3][1], r[4][0], r[4][1], r[5][0], r[5][1], r[6], r[7]))


if __name__ == "__main__":
    main()
