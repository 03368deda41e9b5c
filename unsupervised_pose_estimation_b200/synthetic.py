"""Deterministic synthetic inputs for the view-synthesis loss path.

Lays out the same ``inputs`` / ``outputs`` dictionaries the reference's DataLoader and
networks hand to ``Trainer.generate_images_pred`` / ``compute_losses``
(reference datasets/mono_dataset2.py:129-206 for the key layout, :168-177 for the
per-scale intrinsics, :197-203 for ``stereo_T``; networks/pose_decoder.py:49 for the
0.01 pose scale).  Draw order follows SURVEY.md §8d so numbers are comparable across
runs: colours (frame-major, scale-minor), disparities, then axis-angle / translation.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

K_KITTI = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
K_SCARED = np.array([[0.82, 0, 0.5, 0], [0, 1.02, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
K_LUNG = np.array([[0.635, 0, 0.48, 0], [0, 0.634, 0.50, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)

# BASELINE.json configs: name -> (B, H, W, frame_ids, K_norm)
CONFIGS = {
    "C1": dict(batch=12, height=192, width=640, frame_ids=[0, -1, 1], K=K_KITTI),
    "C2": dict(batch=12, height=256, width=320, frame_ids=[0, -1, 1], K=K_SCARED),
    "C3": dict(batch=12, height=192, width=640, frame_ids=[0, -1, 1, "s"], K=K_KITTI),
    "C4": dict(batch=12, height=320, width=1024, frame_ids=[0, -1, 1], K=K_KITTI),
    "C5": dict(batch=96, height=192, width=640, frame_ids=[0, -1, 1], K=K_KITTI),
}


def scaled_intrinsics(K_norm, height, width, num_scales, batch):
    """Per-scale K and pinv(K), repeated over the batch (mono_dataset2.py:168-177)."""
    out = {}
    for s in range(num_scales):
        K = K_norm.copy()
        K[0, :] *= width // (2 ** s)
        K[1, :] *= height // (2 ** s)
        inv_K = np.linalg.pinv(K)
        out[("K", s)] = torch.from_numpy(K).unsqueeze(0).repeat(batch, 1, 1)
        out[("inv_K", s)] = torch.from_numpy(inv_K).unsqueeze(0).repeat(batch, 1, 1)
    return out


def _draw(gen, shape, family, squash=False):
    if family == "iid":
        return torch.rand(*shape, generator=gen)
    # "smooth": same generator, drawn at 1/8 resolution and bilinearly up-sampled, which gives
    # the gather locality of real images / depth maps.
    b, c, h, w = shape
    lo = torch.rand(b, c, max(h // 8, 2), max(w // 8, 2), generator=gen)
    if squash:
        lo = torch.sigmoid(4.0 * (lo - 0.5))
    return F.interpolate(lo, [h, w], mode="bilinear", align_corners=True).contiguous()


def make_batch(batch, height, width, frame_ids, K_norm=K_KITTI, scales=(0, 1, 2, 3), seed=0,
               family="iid", pose_fn=None, device="cpu", requires_grad=True):
    """Return (inputs, outputs, leaves).

    ``leaves`` holds the tensors gradients are taken with respect to:
    ``disp`` per scale and ``axisangle`` / ``translation`` per temporal frame.
    ``pose_fn(axisangle[B,1,3], translation[B,1,3], invert) -> [B,4,4]`` builds
    ``cam_T_cam`` (the caller passes the implementation under test or the oracle's).
    """
    gen = torch.Generator().manual_seed(seed)
    inputs, outputs, leaves = {}, {}, {}
    for f in frame_ids:
        for s in scales:
            inputs[("color", f, s)] = _draw(gen, (batch, 3, height // 2 ** s, width // 2 ** s), family)
    for s in scales:
        d = _draw(gen, (batch, 1, height // 2 ** s, width // 2 ** s), family, squash=True)
        leaves[("disp", s)] = d
    for f in frame_ids[1:]:
        if f == "s":
            continue
        aa = 0.01 * torch.randn(batch, 2, 1, 3, generator=gen)
        tr = 0.01 * torch.randn(batch, 2, 1, 3, generator=gen)
        leaves[("axisangle", 0, f)] = aa
        leaves[("translation", 0, f)] = tr
    inputs.update(scaled_intrinsics(K_norm, height, width, len(scales), batch))
    if "s" in frame_ids:
        T = torch.eye(4).unsqueeze(0).repeat(batch, 1, 1)
        T[:, 0, 3] = 0.1
        inputs["stereo_T"] = T

    inputs = {k: v.to(device) for k, v in inputs.items()}
    for k in list(leaves):
        leaves[k] = leaves[k].to(device).requires_grad_(requires_grad)
    for s in scales:
        outputs[("disp", s)] = leaves[("disp", s)]
    for f in frame_ids[1:]:
        if f == "s":
            continue
        outputs[("axisangle", 0, f)] = leaves[("axisangle", 0, f)]
        outputs[("translation", 0, f)] = leaves[("translation", 0, f)]
        if pose_fn is not None:
            outputs[("cam_T_cam", 0, f)] = pose_fn(
                leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
    return inputs, outputs, leaves


def make_config(name, **kw):
    cfg = dict(CONFIGS[name])
    K = cfg.pop("K")
    cfg.update(kw)
    return make_batch(cfg.pop("batch"), cfg.pop("height"), cfg.pop("width"), cfg.pop("frame_ids"), K, **cfg)
