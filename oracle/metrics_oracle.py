"""TEST INFRASTRUCTURE -- CPU/PyTorch restatement of the reference's depth metrics and scale-invariant log loss
(the losses / metrics adjacent to the view-synthesis path).  Only tests/, smoke() and
bench.py's baseline legs may import this; the product never does.

Pinned: tests/golden/metrics/metrics.npz holds the outputs of the UNMODIFIED reference functions
(/root/reference/layers.py `SLlog`, `compute_depth_errors`; trainer.py `Trainer.compute_depth_losses`) on seeded
inputs, written by tests/golden/metrics/make_golden_metrics.py; tests/test_metrics_oracle.py requires this file to
reproduce them.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def sllog(fake1, real1):
    """layers.py:32-56 `SLlog.forward` (shapes must match: the reference's resize branch reads an unassigned name)."""
    real = real1.clone()
    fake = fake1.clone()
    N = (real > 0).float().sum()                       # layers.py:44
    mask = ((real <= 0) + (fake <= 0)) > 0              # layers.py:45-49
    fake[mask] = 1.
    real[mask] = 1.
    loss_ = torch.log(real) - torch.log(fake)          # layers.py:53
    return torch.sqrt((torch.sum(loss_ ** 2) / N) - ((torch.sum(loss_) / N) ** 2))   # layers.py:54


def compute_depth_errors(gt, pred):
    """layers.py:335-353: abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3."""
    thresh = torch.max((gt / pred), (pred / gt))
    a1 = (thresh < 1.25).float().mean()
    a2 = (thresh < 1.25 ** 2).float().mean()
    a3 = (thresh < 1.25 ** 3).float().mean()
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(gt - pred) / gt)
    sq_rel = torch.mean((gt - pred) ** 2 / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


def compute_depth_losses(depth_pred, depth_gt):
    """trainer.py:688-716 `Trainer.compute_depth_losses`: up-sample to the KITTI ground-truth size, Garg/Eigen
    crop, median scaling, clamp, then compute_depth_errors.  depth_pred = outputs[("depth", 0, 0)]."""
    depth_pred = torch.clamp(F.interpolate(depth_pred, [375, 1242], mode="bilinear", align_corners=False), 1e-3, 80)
    depth_pred = depth_pred.detach()
    mask = depth_gt > 0
    crop_mask = torch.zeros_like(mask)
    crop_mask[:, :, 153:371, 44:1197] = 1
    mask = mask * crop_mask
    gt = depth_gt[mask]
    pred = depth_pred[mask]
    pred = pred * (torch.median(gt) / torch.median(pred))
    pred = torch.clamp(pred, min=1e-3, max=80)
    return compute_depth_errors(gt, pred)


def depth_to_disp(depth, min_disp=0.00001, max_disp=1.000001):
    """layers.py:74-83."""
    min_depth = 1 / max_disp
    max_depth = 1 / min_disp
    scaled_depth = min_depth + (max_depth - min_depth) * depth
    return scaled_depth, 1 / scaled_depth


def metric_inputs(seed, kind):
    """Seeded inputs shared by the golden generator and the tests (so the fixture stores outputs only)."""
    gen = torch.Generator().manual_seed(seed)
    if kind == "sllog":
        real = torch.rand(2, 1, 24, 40, generator=gen) * 2 - 0.2      # some entries <= 0: masked
        fake = torch.rand(2, 1, 24, 40, generator=gen) * 2 - 0.1
        return fake, real
    if kind == "errors":
        gt = 0.5 + 60 * torch.rand(5000, generator=gen)
        pred = gt * torch.exp(0.3 * torch.randn(5000, generator=gen))
        return gt, pred
    if kind == "depth_losses":
        pred = 1.0 + 40 * torch.rand(2, 1, 48, 160, generator=gen)
        gt = 80 * torch.rand(2, 1, 375, 1242, generator=gen)
        gt = gt * (torch.rand(2, 1, 375, 1242, generator=gen) < 0.05)   # sparse LiDAR-like ground truth
        return pred, gt
    raise ValueError(kind)
