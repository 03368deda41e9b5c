"""ORACLE — test infrastructure, not product code.

A plain-PyTorch restatement of the reference's view-synthesis loss path
(meghakalia/unsupervised_pose_estimation, a monodepth2 fork).  It exists only to
CHECK the CUDA path: nothing under ``unsupervised_pose_estimation_b200/`` imports it.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may call it.

Every function cites the reference lines it restates.  The op sequence (one torch op
per arithmetic step, same operand order) is kept deliberately, because parity of the
discrete decisions (bilinear tap indices, auto-mask arg-min) is judged bit-for-bit and
PyTorch rounds after every op.  The functions are device agnostic: on CPU they are the
``--no_cuda`` reference path, on ``cuda`` they are the eager PyTorch-CUDA reference.

Pinning: the reference ships no tests or golden vectors, so this file is
pinned against the reference ITSELF: ``tests/golden/make_golden.py`` imports
``/root/reference`` (layers.py + the three Trainer methods bound on a namespace), runs
it on seeded inputs and commits the outputs; ``tests/test_oracle_golden.py`` requires
this restatement to reproduce them (bit-exact on CPU for the same torch build).
"""
from __future__ import annotations

import types

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# options
# --------------------------------------------------------------------------------------

def make_opt(**kw):
    """Namespace with the option fields the path reads (reference options.py:59-179).

    Defaults are the reference's own defaults except where SURVEY.md §8d names the
    monodepth2 values used by BASELINE config 1 (max_depth 100, smoothness 1e-3).
    """
    opt = types.SimpleNamespace(
        height=192, width=640, batch_size=12,
        scales=[0, 1, 2, 3], frame_ids=[0, -1, 1],
        min_depth=0.1, max_depth=100.0, disparity_smoothness=1e-3,
        v1_multiscale=False, avg_reprojection=False, disable_automasking=False,
        predictive_mask=False, no_ssim=False, pose_model_type="separate_resnet",
        pre_trained_generator=False,
    )
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt


# --------------------------------------------------------------------------------------
# geometry  (reference layers.py:85-172, 210-264)
# --------------------------------------------------------------------------------------

def disp_to_depth(disp, min_depth, max_depth):
    """layers.py:85-94 — sigmoid disparity -> (scaled disparity, depth)."""
    lo = 1 / max_depth
    hi = 1 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1 / scaled


def rot_from_axisangle(vec):
    """layers.py:133-172 — Rodrigues formula, vec [B,1,3] -> [B,4,4]."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x = axis[..., 0].unsqueeze(1)
    y = axis[..., 1].unsqueeze(1)
    z = axis[..., 2].unsqueeze(1)
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rot = torch.zeros((vec.shape[0], 4, 4), dtype=vec.dtype).to(device=vec.device)  # dtype: fp64 runs of the oracle
    entries = {
        (0, 0): x * xC + ca, (0, 1): xyC - zs, (0, 2): zxC + ys,
        (1, 0): xyC + zs, (1, 1): y * yC + ca, (1, 2): yzC - xs,
        (2, 0): zxC - ys, (2, 1): yzC + xs, (2, 2): z * zC + ca,
    }
    for (r, c), v in entries.items():
        rot[:, r, c] = torch.squeeze(v)
    rot[:, 3, 3] = 1
    return rot


def get_translation_matrix(t):
    """layers.py:117-130 — [B,1,3] -> homogeneous translation [B,4,4]."""
    T = torch.zeros(t.shape[0], 4, 4, dtype=t.dtype).to(device=t.device)
    col = t.contiguous().view(-1, 3, 1)
    for i in range(4):
        T[:, i, i] = 1
    T[:, :3, 3, None] = col
    return T


def transformation_from_parameters(axisangle, translation, invert=False):
    """layers.py:97-114 — pose-net output -> 4x4; inverse is R^T · T(-t)."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t *= -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def pixel_grid(batch, height, width, device, dtype=torch.float32):
    """layers.py:220-232 — homogeneous pixel coordinates [B,3,HW], rows (x, y, 1)."""
    ys, xs = torch.meshgrid(
        torch.arange(height, dtype=dtype), torch.arange(width, dtype=dtype), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, dtype=dtype)], 0)
    return pix.unsqueeze(0).repeat(batch, 1, 1).to(device)


def backproject(depth, inv_K, pix=None):
    """layers.py:234-239 — depth [B,1,h,w], inv_K [B,4,4] -> camera points [B,4,hw]."""
    b, _, h, w = depth.shape
    if pix is None:
        pix = pixel_grid(b, h, w, depth.device, depth.dtype)
    rays = torch.matmul(inv_K[:, :3, :3], pix)
    cam = depth.view(b, 1, -1) * rays
    ones = torch.ones(b, 1, h * w, device=depth.device, dtype=depth.dtype)
    return torch.cat([cam, ones], 1)


def project(points, K, T, height, width, eps=1e-7):
    """layers.py:253-264 — camera points -> normalised sampling grid [B,h,w,2]."""
    b = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, points)
    pix = cam[:, :2, :] / (cam[:, 2, :].unsqueeze(1) + eps)
    pix = pix.view(b, 2, height, width).permute(0, 2, 3, 1)
    pix[..., 0] /= width - 1
    pix[..., 1] /= height - 1
    return (pix - 0.5) * 2


# --------------------------------------------------------------------------------------
# photometric terms  (reference layers.py:286-332, trainer.py:543-555)
# --------------------------------------------------------------------------------------

_C1 = 0.01 ** 2
_C2 = 0.03 ** 2


def ssim(x, y):
    """layers.py:318-332 — per-channel SSIM dissimilarity, 3x3 mean filter, reflect pad 1."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    num = (2 * mu_x * mu_y + _C1) * (2 * sigma_xy + _C2)
    den = (mu_x ** 2 + mu_y ** 2 + _C1) * (sigma_x + sigma_y + _C2)
    return torch.clamp((1 - num / den) / 2, 0, 1)


def reprojection_loss(pred, target, no_ssim=False):
    """trainer.py:543-555 — 0.85·mean_c SSIM + 0.15·mean_c L1 -> [B,1,H,W]."""
    l1 = torch.abs(target - pred).mean(1, True)
    if no_ssim:
        return l1
    return 0.85 * ssim(pred, target).mean(1, True) + 0.15 * l1


def smooth_loss(disp, img):
    """layers.py:286-299 — edge-aware first-order smoothness of a disparity map."""
    dx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
    dy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
    ix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
    iy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
    dx *= torch.exp(-ix)
    dy *= torch.exp(-iy)
    return dx.mean() + dy.mean()


# --------------------------------------------------------------------------------------
# the two Trainer methods  (reference trainer.py:491-541, 557-686)
# --------------------------------------------------------------------------------------

def generate_images_pred(opt, inputs, outputs):
    """trainer.py:491-541 — warp every source frame into the target view per scale.

    Writes ("depth",0,s), ("sample",f,s), ("color",f,s), ("color_identity",f,s).
    """
    for scale in opt.scales:
        disp = outputs[("disp", scale)]
        if opt.v1_multiscale:
            src_scale = scale
        else:
            disp = F.interpolate(disp, [opt.height, opt.width], mode="bilinear", align_corners=False)
            src_scale = 0
        _, depth = disp_to_depth(disp, opt.min_depth, opt.max_depth)
        outputs[("depth", 0, scale)] = depth
        h, w = depth.shape[2], depth.shape[3]
        for frame_id in opt.frame_ids[1:]:
            T = inputs["stereo_T"] if frame_id == "s" else outputs[("cam_T_cam", 0, frame_id)]
            if opt.pose_model_type == "posecnn":  # trainer.py:516-525
                aa = outputs[("axisangle", 0, frame_id)]
                tr = outputs[("translation", 0, frame_id)]
                mean_inv = (1 / depth).mean(3, True).mean(2, True)
                T = transformation_from_parameters(aa[:, 0], tr[:, 0] * mean_inv[:, 0], frame_id < 0)
            cam = backproject(depth, inputs[("inv_K", src_scale)])
            grid = project(cam, inputs[("K", src_scale)], T, h, w)
            outputs[("sample", frame_id, scale)] = grid
            outputs[("color", frame_id, scale)] = F.grid_sample(
                inputs[("color", frame_id, src_scale)], grid,
                padding_mode="border", align_corners=True)
            if not opt.disable_automasking:
                outputs[("color_identity", frame_id, scale)] = inputs[("color", frame_id, src_scale)]


def compute_losses(opt, inputs, outputs, noise=None, generator=None, gen_transform=None):
    """trainer.py:557-686.  The GAN-prior term (:565-583, :684) is evaluated only with --pre_trained_generator;
    ``generator`` / ``gen_transform`` then stand for self.models["pre_trained_generator"] / self.gen_transform.

    ``noise``: optional list (one [B,F,H,W] tensor per scale) replacing the reference's
    ``torch.randn`` draw at trainer.py:656-657; ``None`` draws from the global RNG exactly
    like the reference does.
    """
    losses = {}
    total = 0
    gan_total = 0
    n_src = len(opt.frame_ids) - 1
    if opt.pre_trained_generator:  # trainer.py:565-583
        from oracle.metrics_oracle import depth_to_disp, sllog
        fake_B1 = generator(gen_transform(inputs[("color", 0, 0)]))
        _, fake_disp_scaled = depth_to_disp(fake_B1)
        for scale in opt.scales:
            disp = F.interpolate(outputs[("disp", scale)], [opt.height, opt.width], mode="bilinear", align_corners=False)
            gan_loss = sllog(fake_disp_scaled, disp)
            losses["gan_loss/{}".format(scale)] = gan_loss
            gan_total = gan_total + gan_loss
    for si, scale in enumerate(opt.scales):
        loss = 0
        src_scale = scale if opt.v1_multiscale else 0
        disp = outputs[("disp", scale)]
        color = inputs[("color", 0, scale)]
        target = inputs[("color", 0, src_scale)]

        reproj = torch.cat(
            [reprojection_loss(outputs[("color", f, scale)], target, opt.no_ssim)
             for f in opt.frame_ids[1:]], 1)

        if not opt.disable_automasking:
            ident = torch.cat(
                [reprojection_loss(inputs[("color", f, src_scale)], target, opt.no_ssim)
                 for f in opt.frame_ids[1:]], 1)
            if opt.avg_reprojection:
                ident = ident.mean(1, keepdim=True)
        elif opt.predictive_mask:  # trainer.py:635-647 (device-agnostic instead of the hard-coded .cuda())
            mask = outputs["predictive_mask"][("disp", scale)]
            if not opt.v1_multiscale:
                mask = F.interpolate(mask, [opt.height, opt.width], mode="bilinear", align_corners=False)
            reproj *= mask
            weighting = 0.2 * torch.nn.BCELoss()(mask, torch.ones(mask.shape, device=mask.device))
            loss += weighting.mean()
        if opt.avg_reprojection:
            reproj = reproj.mean(1, keepdim=True)

        if not opt.disable_automasking:
            if noise is None:
                z = torch.randn(ident.shape, device=ident.device)
            else:
                z = noise[si]
            ident += z * 0.00001
            combined = torch.cat((ident, reproj), dim=1)
        else:
            combined = reproj

        if combined.shape[1] == 1:
            to_opt = combined
        else:
            to_opt, idxs = torch.min(combined, dim=1)
        if not opt.disable_automasking:
            outputs["identity_selection/{}".format(scale)] = (idxs > ident.shape[1] - 1).float()

        loss += to_opt.mean()
        losses["min_loss/{}".format(scale)] = to_opt.mean()

        mean_disp = disp.mean(2, True).mean(3, True)
        norm_disp = disp / (mean_disp + 1e-7)
        loss += opt.disparity_smoothness * smooth_loss(norm_disp, color) / (2 ** scale)
        total += loss
        losses["loss/{}".format(scale)] = loss
    total /= len(opt.scales)
    losses["loss"] = total + gan_total / len(opt.scales) * 0.002   # trainer.py:684
    assert n_src >= 1
    return losses


def loss_step(opt, inputs, outputs, noise=None):
    """One pass of the path: warp + losses (the forward half of a 'step')."""
    generate_images_pred(opt, inputs, outputs)
    return compute_losses(opt, inputs, outputs, noise)
