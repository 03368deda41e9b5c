"""Drop-in for the reference ``layers.py`` — same names, constructor arguments and call semantics —
with the view-synthesis layers executing in hand-written sm_100a kernels (libvsl_b200.so).

CUDA-backed (the hot path; reference layers.py lines in brackets):
    disp_to_depth [85-94] (thin torch arithmetic, identical op order), BackprojectDepth [210-239],
    Project3D [242-264], get_smooth_loss [286-299], SSIM [302-332].
Pose helpers: transformation_from_parameters [97-114] is one CUDA kernel
for CUDA inputs (bit-identical to the torch ops, with an analytic backward) and the original torch ops
otherwise; get_translation_matrix [117-130] and rot_from_axisangle [133-172] stay plain torch.
Adjacent losses / metrics: SLlog [32-56] and compute_depth_errors [335-353] run as
CUDA kernels for CUDA inputs (fixed-order fp64 reductions), the reference's torch ops otherwise.
Names re-exported only so ``from layers import *`` users keep working (networks/depth_decoder.py:14,
evaluate_depth.py:10): RMSE_log, depth_to_disp, ConvBlock, Conv3x3, batchNorm, upsample, deconv.  They are
network blocks, not part of the path.
"""
from __future__ import absolute_import, division, print_function

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as VF

__all__ = [
    "SLlog", "RMSE_log", "depth_to_disp", "disp_to_depth", "transformation_from_parameters",
    "get_translation_matrix", "rot_from_axisangle", "ConvBlock", "batchNorm", "Conv3x3",
    "BackprojectDepth", "Project3D", "upsample", "deconv", "get_smooth_loss", "SSIM",
    "compute_depth_errors",
]


# ----------------------------------------------------------------------------------------------
# the path
# ----------------------------------------------------------------------------------------------
def disp_to_depth(disp, min_depth, max_depth):
    """Sigmoid disparity -> (scaled disparity, depth); reference layers.py:85-94."""
    lo, hi = 1 / max_depth, 1 / min_depth
    scaled_disp = lo + (hi - lo) * disp
    return scaled_disp, 1 / scaled_disp


class BackprojectDepth(nn.Module):
    """Depth image -> homogeneous camera points [B,4,h*w] (reference layers.py:210-239).

    The reference keeps three constant buffers (pixel grid, ones, homogeneous grid: 23.6 MB at
    640x192x12); here the pixel coordinates come from the thread index, so the module is stateless.
    ``batch_size/height/width`` are kept and validated like the reference's ``view`` would.
    """

    def __init__(self, batch_size, height, width):
        super(BackprojectDepth, self).__init__()
        self.batch_size, self.height, self.width = batch_size, height, width

    def forward(self, depth, inv_K):
        if depth.numel() != self.batch_size * self.height * self.width:
            raise RuntimeError("BackprojectDepth(%d,%d,%d) got depth of shape %s" % (
                self.batch_size, self.height, self.width, tuple(depth.shape)))
        return VF.backproject(depth.reshape(self.batch_size, 1, self.height, self.width), inv_K)


class Project3D(nn.Module):
    """Camera points -> normalised sampling grid [B,h,w,2] for a camera (K, T)
    (reference layers.py:242-264).  K@T is 16 floats per image and stays a torch.matmul so autograd
    reaches the pose network; the per-pixel projection is the CUDA kernel."""

    def __init__(self, batch_size, height, width, eps=1e-7):
        super(Project3D, self).__init__()
        self.batch_size, self.height, self.width, self.eps = batch_size, height, width, eps

    def forward(self, points, K, T):
        P = torch.matmul(K, T)[:, :3, :]
        if points.shape[0] != self.batch_size or points.shape[-1] != self.height * self.width:
            raise RuntimeError("Project3D(%d,%d,%d) got points of shape %s" % (
                self.batch_size, self.height, self.width, tuple(points.shape)))
        return VF.project(points, P, self.height, self.width, self.eps)


def get_smooth_loss(disp, img):
    """Edge-aware disparity smoothness (reference layers.py:286-299); differentiable w.r.t. disp."""
    return VF.smooth_loss(disp, img)


class SSIM(nn.Module):
    """Per-channel SSIM dissimilarity, 3x3 mean filter over a reflection-padded image
    (reference layers.py:302-332)."""

    def __init__(self):
        super(SSIM, self).__init__()
        self.C1 = 0.01 ** 2
        self.C2 = 0.03 ** 2

    def forward(self, x, y):
        return VF.ssim(x, y)


# ----------------------------------------------------------------------------------------------
# pose helpers (host-side torch; gradients flow through them to the pose network)
# ----------------------------------------------------------------------------------------------
def rot_from_axisangle(vec):
    """Axis-angle [B,1,3] -> rotation as 4x4 (reference layers.py:133-172, Rodrigues)."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rows = [
        [x * xC + ca, xyC - zs, zxC + ys],
        [xyC + zs, y * yC + ca, yzC - xs],
        [zxC - ys, yzC + xs, z * zC + ca],
    ]
    rot = torch.zeros((vec.shape[0], 4, 4)).to(device=vec.device)
    for r in range(3):
        for c in range(3):
            rot[:, r, c] = torch.squeeze(rows[r][c])
    rot[:, 3, 3] = 1
    return rot


def get_translation_matrix(translation_vector):
    """[B,1,3] -> homogeneous translation (reference layers.py:117-130)."""
    T = torch.zeros(translation_vector.shape[0], 4, 4).to(device=translation_vector.device)
    t = translation_vector.contiguous().view(-1, 3, 1)
    for i in range(4):
        T[:, i, i] = 1
    T[:, :3, 3, None] = t
    return T


def transformation_from_parameters(axisangle, translation, invert=False):
    """Pose-net (axisangle, translation) -> 4x4 (reference layers.py:97-114).

    CUDA float32 tensors go through one kernel (vsl_pose_forward / _backward) that reproduces the torch op
    sequence below bit for bit; anything else (the CPU callers evaluate_pose.py / test_simple.py, float64)
    runs the original torch ops, which are this helper's definition."""
    if axisangle.is_cuda and axisangle.dtype == torch.float32 and translation.dtype == torch.float32 \
            and axisangle.dim() == 3 and axisangle.shape[1:] == (1, 3) and translation.shape == axisangle.shape:
        return VF.pose_matrix(axisangle, translation, invert)
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t *= -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


# ----------------------------------------------------------------------------------------------
# re-exports outside the path (plain PyTorch)
# ----------------------------------------------------------------------------------------------
def depth_to_disp(depth, min_disp=0.00001, max_disp=1.000001):
    """Inverse mapping used by the GAN prior branch (reference layers.py:74-83)."""
    lo, hi = 1 / max_disp, 1 / min_disp
    scaled_depth = lo + (hi - lo) * depth
    return scaled_depth, 1 / scaled_depth


def _valid_log_pair(fake, real):
    real, fake = real.clone(), fake.clone()
    bad = (real <= 0) | (fake <= 0)
    real[bad] = 1.
    fake[bad] = 1.
    return torch.log(real) - torch.log(fake)


class SLlog(nn.Module):
    """Scale-invariant log loss of the GAN-prior branch (reference layers.py:32-56)."""

    def forward(self, fake1, real1):
        if fake1.is_cuda and fake1.dtype == torch.float32:
            return VF.sllog(fake1, real1)   # one reduction kernel forward, one element-wise kernel backward
        n = (real1 > 0).float().sum()
        d = _valid_log_pair(fake1, real1)
        return torch.sqrt((torch.sum(d ** 2) / n) - ((torch.sum(d) / n) ** 2))


class RMSE_log(nn.Module):
    """Log-RMSE on pixels with real < 1 (reference layers.py:58-72)."""

    def __init__(self, use_cuda):
        super(RMSE_log, self).__init__()
        self.eps = 1e-8
        self.use_cuda = use_cuda

    def forward(self, fake, real):
        sel = real < 1.
        fake = F.interpolate(fake, size=real.shape[2:], mode="bilinear") + self.eps
        d = torch.log(real[sel]) - torch.log(fake[sel])
        return torch.sqrt(torch.sum(torch.abs(d) ** 2) / d.numel())


class Conv3x3(nn.Module):
    """3x3 convolution after reflection (or zero) padding (reference layers.py:192-207)."""

    def __init__(self, in_channels, out_channels, use_refl=True):
        super(Conv3x3, self).__init__()
        self.pad = nn.ReflectionPad2d(1) if use_refl else nn.ZeroPad2d(1)
        self.conv = nn.Conv2d(int(in_channels), int(out_channels), 3)

    def forward(self, x):
        return self.conv(self.pad(x))


class ConvBlock(nn.Module):
    """Conv3x3 + ELU (reference layers.py:175-187)."""

    def __init__(self, in_channels, out_channels):
        super(ConvBlock, self).__init__()
        self.conv = Conv3x3(in_channels, out_channels)
        self.nonlin = nn.ELU(inplace=True)

    def forward(self, x):
        return self.nonlin(self.conv(x))


def batchNorm(num_ch_dec):
    return nn.BatchNorm2d(num_ch_dec)


def upsample(x):
    """Nearest-neighbour x2 (reference layers.py:267-270)."""
    return F.interpolate(x, scale_factor=2, mode="nearest")


class deconv(nn.Module):
    """Stride-2 transposed convolution (reference layers.py:272-282)."""

    def __init__(self, ch_in, ch_out):
        super(deconv, self).__init__()
        self.deconvlayer = nn.ConvTranspose2d(ch_in, ch_out, 3, stride=2, padding=1)

    def forward(self, x):
        return self.deconvlayer(x)


def compute_depth_errors(gt, pred):
    """KITTI depth metrics (reference layers.py:335-353).  CUDA fp32 inputs: one kernel (VF.depth_errors); the
    seven results are 0-dim views of one device vector, like the reference's seven 0-dim tensors."""
    if gt.is_cuda and gt.dtype == torch.float32 and pred.dtype == torch.float32:
        return tuple(VF.depth_errors(gt, pred).unbind(0))
    ratio = torch.max(gt / pred, pred / gt)
    a1, a2, a3 = ((ratio < 1.25 ** k).float().mean() for k in (1, 2, 3))
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(gt - pred) / gt)
    sq_rel = torch.mean((gt - pred) ** 2 / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3
