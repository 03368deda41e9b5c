"""GPU tests of the losses / metrics next to the path: SLlog, compute_depth_errors,
Trainer.compute_depth_losses and the --pre_trained_generator term, CUDA kernels against the oracle
(oracle/metrics_oracle.py, pinned to the reference by tests/golden/metrics) and the committed goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M
from oracle import vsl_oracle as O
from unsupervised_pose_estimation_b200 import layers as L
from unsupervised_pose_estimation_b200 import functional as VF
from unsupervised_pose_estimation_b200 import synthetic
from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics", "metrics.npz"))


@pytest.mark.parametrize("seed", [0, 1])
def test_sllog_kernel(seed):
    fake, real = (t.to(DEV).requires_grad_(True) for t in M.metric_inputs(seed, "sllog"))
    loss = L.SLlog()(fake, real)
    gf, gr = torch.autograd.grad(loss * 1.7, [fake, real])
    ref = M.sllog(fake, real)
    rf, rr = torch.autograd.grad(ref * 1.7, [fake, real])
    assert abs(loss.item() - ref.item()) <= 1e-6 * ref.item()
    assert abs(loss.item() - float(GOLD["sllog|%d|loss" % seed])) <= 2e-6 * ref.item()   # the reference's own value
    for a, b in ((gf, rf), (gr, rr)):
        assert ((a - b).norm() / b.norm()).item() <= 1e-5
    assert np.allclose(gf.cpu().numpy() / 1.7, GOLD["sllog|%d|grad_fake" % seed], rtol=2e-4, atol=1e-8)
    # masked entries (real <= 0 or fake <= 0) carry no gradient
    dead = (real <= 0) | (fake <= 0)
    assert dead.any() and float(gf[dead].abs().max()) == 0.0 and float(gr[dead].abs().max()) == 0.0


@pytest.mark.parametrize("seed", [0, 1])
def test_compute_depth_errors_kernel(seed):
    gt, pred = (t.to(DEV) for t in M.metric_inputs(seed, "errors"))
    got = L.compute_depth_errors(gt, pred)
    assert len(got) == 7 and all(v.dim() == 0 for v in got)
    ref = M.compute_depth_errors(gt, pred)
    for a, b, g in zip(got, ref, GOLD["errors|%d" % seed]):
        assert abs(a.item() - b.item()) <= 2e-6 * abs(b.item())
        assert abs(a.item() - g) <= 1e-5 * abs(g)
    # a big ragged size: more elements than one pass of the grid
    gen = torch.Generator().manual_seed(3)
    gt = (0.5 + 70 * torch.rand(1_234_567, generator=gen)).to(DEV)
    pred = gt * torch.exp(0.2 * torch.randn(1_234_567, generator=gen).to(DEV))
    for a, b in zip(L.compute_depth_errors(gt, pred), M.compute_depth_errors(gt, pred)):
        assert abs(a.item() - b.item()) <= 1e-5 * abs(b.item())
    assert torch.equal(VF.depth_errors(gt, pred), VF.depth_errors(gt, pred))   # fixed-order reduction


@pytest.mark.parametrize("seed", [0, 1])
def test_compute_depth_losses_kernels(seed):
    """trainer.py:688-716: up-sample + crop + exact median scaling + metrics."""
    pred, gt = (t.to(DEV) for t in M.metric_inputs(seed, "depth_losses"))
    path = LossPath(make_opt(height=48, width=160, batch_size=2), device=DEV, side_outputs="none")
    losses = {}
    path.compute_depth_losses({"depth_gt": gt}, {("depth", 0, 0): pred}, losses)
    ref = M.compute_depth_losses(pred, gt)
    assert list(losses) == path.depth_metric_names
    for k, b, g in zip(path.depth_metric_names, ref, GOLD["depth_losses|%d" % seed]):
        assert isinstance(losses[k], np.ndarray) and losses[k].shape == ()
        assert abs(float(losses[k]) - b.item()) <= 1e-5 * abs(b.item()), k
        assert abs(float(losses[k]) - g) <= 1e-4 * abs(g), k
    # the median scaling itself is exact: recompute it from the kernel's own masked values
    up = torch.clamp(torch.nn.functional.interpolate(pred, [375, 1242], mode="bilinear", align_corners=False), 1e-3, 80)
    mask = gt > 0
    crop = torch.zeros_like(mask)
    crop[:, :, 153:371, 44:1197] = 1
    mask = mask & crop
    # an even and an odd number of valid pixels both take the LOWER median (torch.median)
    for drop in (0, 1):
        m2 = mask.clone()
        if drop:
            idx = m2.nonzero()[0]
            m2[tuple(idx)] = False
        gt2 = gt * m2
        out = VF.depth_losses(pred, gt2)
        r = M.compute_depth_losses(pred, gt2)
        for a, b in zip(out, r):
            assert abs(a.item() - b.item()) <= 1e-5 * abs(b.item())
    assert up.shape[-2:] == (375, 1242)


def test_pre_trained_generator_term():
    """--pre_trained_generator (trainer.py:565-583, :684): gan_loss/s entries and their share of the total, with
    gradients reaching the disparities through the SLlog kernel."""
    B, H, W, frames = 2, 64, 96, [0, -1, 1]
    torch.manual_seed(0)
    gen = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(4, 1, 3, padding=1),
                              torch.nn.Sigmoid()).to(DEV)
    gray = lambda img: img.mean(1, keepdim=True)
    opt = O.make_opt(height=H, width=W, batch_size=B, frame_ids=frames, pre_trained_generator=True)
    inputs, outputs, leaves = synthetic.make_batch(B, H, W, frames, seed=12, family="smooth", device=DEV)

    def poses(fn):
        out = dict(outputs)
        for f in frames[1:]:
            out[("cam_T_cam", 0, f)] = fn(leaves[("axisangle", 0, f)][:, 0], leaves[("translation", 0, f)][:, 0], f < 0)
        return out
    ref_out = poses(O.transformation_from_parameters)
    O.generate_images_pred(opt, inputs, ref_out)
    torch.manual_seed(4)
    ref = O.compute_losses(opt, inputs, ref_out, generator=gen, gen_transform=gray)
    ref_g = torch.autograd.grad(ref["loss"], list(leaves.values()))
    path = LossPath(make_opt(**vars(opt)), device=DEV, side_outputs="none")
    path.models = {"pre_trained_generator": gen}
    path.gen_transform = gray
    out = poses(L.transformation_from_parameters)
    torch.manual_seed(4)
    losses = path.compute_losses(inputs, out)
    g = torch.autograd.grad(losses["loss"], list(leaves.values()))
    assert set(losses) == set(ref) and "gan_loss/3" in losses
    # sqrt(E[d^2] - E[d]^2) cancels: the reference's fp32 sums carry ~1e-5 of noise into it; the kernel accumulates in
    # fp64, so it is compared tightly with the fp64 evaluation of the same formula and loosely with the fp32 one
    fake64 = M.depth_to_disp(gen(gray(inputs[("color", 0, 0)])))[1].double()
    for k in ref:
        if k.startswith("gan_loss"):
            disp64 = torch.nn.functional.interpolate(outputs[("disp", int(k[-1]))], [H, W], mode="bilinear",
                                                     align_corners=False).double()
            exact = M.sllog(fake64, disp64).item()
            assert abs(losses[k].item() - exact) <= 2e-6 * exact, k
            assert abs(losses[k].item() - ref[k].item()) <= 1e-4 * abs(ref[k].item()), k
        else:
            assert abs(losses[k].item() - ref[k].item()) <= 2e-6 * abs(ref[k].item()), k
    for a, b in zip(g, ref_g):
        assert ((a - b).norm() / b.norm()).item() <= 5e-5
    with pytest.raises(RuntimeError):
        LossPath(make_opt(**vars(opt)), device=DEV, side_outputs="none").compute_losses(inputs, poses(L.transformation_from_parameters))
