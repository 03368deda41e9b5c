"""A minimal self-supervised training loop on the B200 loss path (stand-in networks, synthetic frames).

Shows how the pieces of this repository sit in the reference's ``Trainer.process_batch`` / ``run_epoch``
(trainer.py:297-343, :370-403):

    8-bit frames (host, pinned) --H2D--> LossInputPipeline  (pyramid + ToTensor on the GPU, Pillow-exact)
    depth net (stand-in)  -> outputs[("disp", s)]            s = 0..3
    pose net  (stand-in)  -> axis-angle / translation -> transformation_from_parameters -> ("cam_T_cam", 0, f)
    ViewSynthesisLossMixin.generate_images_pred + compute_losses   (one fused CUDA kernel, forward + backward)
    losses["loss"].backward() -> Adam step                   (+ parallel.GradBuckets under torchrun: bucketed
                                                              all-reduce overlapped with the backward)

    python examples/minimal_training.py [--steps 20]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/minimal_training.py

The two networks are a few convolutions each — enough to have parameters to train; the reference's ResNet
encoders/decoders (networks/) plug in at the same two places unchanged.
"""
import argparse
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsupervised_pose_estimation_b200 import parallel, synthetic          # noqa: E402
from unsupervised_pose_estimation_b200.input_pipeline import LossInputPipeline  # noqa: E402
from unsupervised_pose_estimation_b200.layers import transformation_from_parameters  # noqa: E402
from unsupervised_pose_estimation_b200.trainer import ViewSynthesisLossMixin, make_opt  # noqa: E402


class TinyDepthNet(nn.Module):
    """Stand-in for networks/resnet_encoder.py + depth_decoder.py: sigmoid disparities at four scales."""

    def __init__(self, scales):
        super().__init__()
        self.scales = scales
        self.stem = nn.Sequential(nn.Conv2d(3, 16, 3, padding=1), nn.ELU(), nn.Conv2d(16, 16, 3, padding=1), nn.ELU())
        self.heads = nn.ModuleList([nn.Conv2d(16, 1, 3, padding=1) for _ in scales])

    def forward(self, img):
        feat = self.stem(img)
        out = {}
        for s, head in zip(self.scales, self.heads):
            f = feat if s == 0 else F.avg_pool2d(feat, 2 ** s)
            out[("disp", s)] = torch.sigmoid(head(f))
        return out


class TinyPoseNet(nn.Module):
    """Stand-in for the pose encoder/decoder: axis-angle and translation of one frame pair (x 0.01, pose_decoder.py:49)."""

    def __init__(self):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(6, 16, 7, stride=4, padding=3), nn.ReLU(), nn.Conv2d(16, 6, 3, padding=1))

    def forward(self, a, b):
        out = 0.01 * self.net(torch.cat([a, b], 1)).mean((2, 3))
        return out[:, None, :3], out[:, None, 3:]


class MiniTrainer(ViewSynthesisLossMixin):
    """The loss half of the reference Trainer comes from the mixin; everything else is the loop below."""

    def __init__(self, opt, device):
        self.opt, self.device = opt, device
        self.num_scales = len(opt.scales)
        self.vsl_side_outputs = "fused"   # reference-visible outputs at +0.02 ms
        self.depth = TinyDepthNet(opt.scales).to(device)
        self.pose = TinyPoseNet().to(device)
        self.params = list(self.depth.parameters()) + list(self.pose.parameters())
        self.optim = torch.optim.Adam(self.params, 1e-3)
        # .grad of every parameter becomes a view into one flat buffer; hooks all-reduce bucket by bucket during backward
        self.buckets = parallel.GradBuckets(self.params, local_batch=opt.batch_size, bucket_bytes=1 << 16)
        self.inputs_from_frames = LossInputPipeline(opt, device)

    def process_batch(self, inputs):
        outputs = self.depth(inputs[("color", 0, 0)])
        for f in self.opt.frame_ids[1:]:                       # trainer.py:420-438
            pair = (inputs[("color", f, 0)], inputs[("color", 0, 0)]) if f < 0 else (inputs[("color", 0, 0)], inputs[("color", f, 0)])
            axisangle, translation = self.pose(*pair)
            outputs[("cam_T_cam", 0, f)] = transformation_from_parameters(axisangle, translation, invert=(f < 0))  # [B,1,3] each
        self.generate_images_pred(inputs, outputs)             # trainer.py:399
        return outputs, self.compute_losses(inputs, outputs)   # trainer.py:400

    def step(self, frames_u8, intrinsics):
        inputs = dict(intrinsics)
        self.inputs_from_frames({f: v.to(self.device, non_blocking=True) for f, v in frames_u8.items()}, inputs)
        outputs, losses = self.process_batch(inputs)
        self.buckets.begin_step()                              # trainer.py:311 zero_grad, keeping the flat views
        losses["loss"].backward()                              # trainer.py:312; bucket all-reduces start from the hooks
        self.buckets.finish()                                  # no exchange unless torch.distributed is initialised
        self.optim.step()
        return outputs, losses


def synthetic_frames(opt, seed):
    """8-bit HWC frames of a smooth scene seen from three slightly shifted positions, pinned, + intrinsics."""
    inputs, _, _ = synthetic.make_batch(opt.batch_size, opt.height, opt.width, opt.frame_ids, seed=seed, family="smooth",
                                        device="cpu", requires_grad=False)
    base = inputs[("color", 0, 0)]
    frames = {}
    for f in opt.frame_ids:
        img = torch.roll(base, shifts=2 * f, dims=3)           # a camera translation, as far as the loss can tell
        frames[f] = (img.permute(0, 2, 3, 1) * 255).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
    intr = {k: v for k, v in inputs.items() if isinstance(k, tuple) and k[0] in ("K", "inv_K")}
    return frames, intr


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--height", type=int, default=96)
    ap.add_argument("--width", type=int, default=160)
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(0)                                       # same initial weights on every rank
    opt = make_opt(height=args.height, width=args.width, batch_size=args.batch, frame_ids=[0, -1, 1])
    trainer = MiniTrainer(opt, device)
    frames, intr = synthetic_frames(opt, seed=rank)            # every rank its own shard of the data
    intr = {k: v.to(device) for k, v in intr.items()}
    history = []
    for i in range(args.steps):
        outputs, losses = trainer.step(frames, intr)
        logged = parallel.all_reduce_losses(losses, opt.batch_size)
        history.append(float(logged["loss"].detach()))
        if rank == 0 and (i % 5 == 0 or i == args.steps - 1):
            print("step %3d  loss %.5f  auto-masked %.1f %%" % (
                i, history[-1], 100 * (1 - outputs["identity_selection/0"].mean().item())))
    if world > 1:
        dist.destroy_process_group()
    return history


if __name__ == "__main__":
    main()
