"""Developer measurement (not part of bench.py): the reference's op stream run as eager PyTorch-CUDA on
the same GPU (the oracle port on `cuda`) next to the fused path, same inputs, CUDA events.
Writes gpurun_out/eager_cuda.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vsl_oracle as O  # noqa: E402
from unsupervised_pose_estimation_b200 import layers as L, synthetic  # noqa: E402
from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
out = {}
for name in ("C1", "C2", "C4"):
    cfg = dict(synthetic.CONFIGS[name])
    B, H, W, frames = cfg["batch"], cfg["height"], cfg["width"], cfg["frame_ids"]
    opt = O.make_opt(height=H, width=W, batch_size=B, frame_ids=list(frames))
    inputs, outputs, leaves = synthetic.make_batch(B, H, W, frames, cfg["K"], seed=0, family="smooth", device=dev)
    poses = {("cam_T_cam", 0, f): L.transformation_from_parameters(
        leaves[("axisangle", 0, f)][:, 0].detach(), leaves[("translation", 0, f)][:, 0].detach(), f < 0).requires_grad_(True)
        for f in frames[1:] if f != "s"}
    lv = {k: v for k, v in leaves.items() if k[0] == "disp"}
    lv.update(poses)

    def eager():
        o = dict(lv)
        losses = O.loss_step(opt, inputs, o)
        return torch.autograd.grad(losses["loss"], list(lv.values()))

    path = LossPath(make_opt(**vars(opt)), device=dev, side_outputs="none")

    def fused():
        o = dict(lv)
        path.generate_images_pred(inputs, o)
        losses = path.compute_losses(inputs, o)
        return torch.autograd.grad(losses["loss"], list(lv.values()))

    res = {}
    for tag, fn, n in (("eager_torch_cuda", eager, 20), ("fused_eager_launches", fused, 100)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[tag] = {"ms_per_step": ms, "px_per_s": B * H * W / ms * 1e3}
    res["speedup"] = res["eager_torch_cuda"]["ms_per_step"] / res["fused_eager_launches"]["ms_per_step"]
    out[name] = res
    print(name, res)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "eager_cuda.json"), "w"), indent=1)
