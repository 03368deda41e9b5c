"""Drives every product kernel a few times at config C1 so ncu can capture the small ones (tools/gpu_small_kernels.sh):
input pipeline (target: 4 levels, sources: level 0), side outputs (k_warp_forward), fused loss + epilogue, backward
(k_combine), pose kernels, source-image gradient kernels, metrics kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unsupervised_pose_estimation_b200 import functional as VF, layers as L, synthetic  # noqa: E402
from unsupervised_pose_estimation_b200.input_pipeline import ColorAugment, LossInputPipeline, draw_color_aug_params  # noqa: E402
from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
cfg = synthetic.CONFIGS["C1"]
B, H, W, frames = cfg["batch"], cfg["height"], cfg["width"], cfg["frame_ids"]
opt = make_opt(height=H, width=W, batch_size=B, frame_ids=list(frames), max_depth=100.0, disparity_smoothness=1e-3)
inputs, outputs, leaves = synthetic.make_batch(B, H, W, frames, cfg["K"], seed=0, family="smooth", device="cuda")
pipe = LossInputPipeline(opt, "cuda")
u8 = {f: (inputs[("color", f, 0)].permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8).contiguous() for f in frames}
flip = torch.tensor([i % 2 for i in range(B)], dtype=torch.uint8, device="cuda")
path = LossPath(opt, device="cuda", side_outputs="eager")
aug = ColorAugment(B, H, W)
torch.manual_seed(0)
aug_params = [dict(draw_color_aug_params(), autocontrast=True) for _ in range(B)]   # every kernel has work in every image
for it in range(int(os.environ.get("ITERS", "3"))):
    aug(u8[0], aug_params)
    pin = pipe(u8, flip=flip if it == 2 else None)
    pin.update({k: v for k, v in inputs.items() if k[0] in ("K", "inv_K")})
    out = dict(outputs)
    for f in frames[1:]:
        out[("cam_T_cam", 0, f)] = L.transformation_from_parameters(leaves[("axisangle", 0, f)][:, 0],
                                                                   leaves[("translation", 0, f)][:, 0], f < 0)
    path.generate_images_pred(pin, out)
    losses = path.compute_losses(pin, out)
    torch.autograd.grad(losses["loss"], list(leaves.values()))
    # source-image gradients + metrics
    pin2 = dict(pin)
    for f in frames[1:]:
        pin2[("color", f, 0)] = pin[("color", f, 0)].clone().requires_grad_(True)
    l2 = path.compute_losses(pin2, out)
    torch.autograd.grad(l2["loss"], [pin2[("color", f, 0)] for f in frames[1:]])
    L.compute_depth_errors(out[("depth", 0, 0)].reshape(-1), out[("depth", 0, 1)].reshape(-1))
    gt = out[("depth", 0, 0)].detach() * (torch.rand(B, 1, H, W, device="cuda") < 0.05)
    VF.depth_losses(out[("depth", 0, 1)], torch.nn.functional.interpolate(gt, [375, 1242]))
    L.SLlog()(out[("disp", 0)], out[("depth", 0, 0)].detach().clamp(0, 1))
torch.cuda.synchronize()
print("ok")
