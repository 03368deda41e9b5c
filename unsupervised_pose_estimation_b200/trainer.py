"""Drop-in for the loss half of the reference ``Trainer`` (reference trainer.py:491-686).

``ViewSynthesisLossMixin`` provides ``generate_images_pred``, ``compute_reprojection_loss`` and
``compute_losses`` with the reference's signatures; a maintainer mixes it into the reference
``Trainer``.  ``LossPath`` is the same thing stand-alone: it carries exactly the
attributes those methods read (``opt``, ``device``, ``num_scales``), so tests and the benchmark can
drive the path without the rest of the trainer.

How the work is split (differs from the reference on purpose):
  * ``compute_losses`` launches ONE fused forward+backward pass (libvsl_b200.so) that warps, scores,
    auto-masks, reduces and produces the gradients w.r.t. every ``("disp", s)`` and every
    ``P_f = (K @ T_f)[:, :3, :]``; autograd connects those to the depth and pose networks.
  * ``generate_images_pred`` only has to provide the reference's *side outputs*
    (``("depth",0,s)``, ``("sample",f,s)``, ``("color",f,s)``, ``("color_identity",f,s)``), which the
    reference reads on logging steps only (wandb_logging.py:134-143).  ``vsl_side_outputs``:
    ``"eager"`` (default, faithful: written every call by a small CUDA kernel), ``"fused"`` (the tensors
    are put into ``outputs`` by ``generate_images_pred`` and filled by the fused kernel of the following
    ``compute_losses``: same values, a fifth of the cost), ``"none"`` (skip; call
    ``materialize_side_outputs`` on logging steps).
The tie-break noise is drawn with ``torch.randn`` in the reference's order (trainer.py:656-657), so
the global RNG stream is consumed identically.
"""
from __future__ import annotations

import types

import numpy as np
import torch

from . import _lib
from . import functional as VF


def _unsupported(opt):
    bad = []
    if getattr(opt, "pose_model_type", "separate_resnet") == "posecnn" and "s" in opt.frame_ids:
        bad.append("--pose_model_type posecnn with a stereo frame (the reference itself fails there)")
    return bad


class ViewSynthesisLossMixin:
    """Methods of reference ``Trainer`` on the view-synthesis loss path, CUDA-backed."""

    vsl_side_outputs = "eager"
    vsl_arith = "auto"  # VSL_ARITH_* bits, or "auto": calibrate against torch.bmm once per shape
    # --pose_model_type posecnn (trainer.py:516-525): "torch" forms the per-scale poses with the reference's own ops
    # (bit-identical sampling grids); "fused" uses vsl_posecnn_forward/backward — three launches instead of ~40 torch
    # ops per step, the mean inverse depth accumulated in fp64 (poses agree to ~1e-7, not bit for bit)
    vsl_posecnn_tail = "torch"

    # -- plan ---------------------------------------------------------------------------------
    vsl_image_dtype = torch.float32  # storage of the colour images: torch.float32 or torch.bfloat16

    def _vsl_plan(self, image_dtype=None):
        opt = self.opt
        if image_dtype is not None:
            self.vsl_image_dtype = image_dtype
        flags = 0
        if not getattr(opt, "disable_automasking", False):
            flags |= _lib.FLAG_AUTOMASK
        if getattr(opt, "no_ssim", False):
            flags |= _lib.FLAG_NO_SSIM
        if getattr(opt, "avg_reprojection", False):
            flags |= _lib.FLAG_AVG_REPROJECTION
        v1 = bool(getattr(opt, "v1_multiscale", False))
        key = (opt.batch_size, opt.height, opt.width, tuple(opt.scales), len(opt.frame_ids) - 1,
               opt.min_depth, opt.max_depth, opt.disparity_smoothness, self.vsl_arith, flags, self.vsl_image_dtype, v1)
        plan = getattr(self, "_vsl_plan_cache", None)
        if plan is None or plan[0] != key:
            bad = _unsupported(opt)
            if bad:
                raise NotImplementedError(
                    "the CUDA view-synthesis path implements the reference's default loss "
                    "(automask + per-pixel min + SSIM at full resolution); not yet: " + ", ".join(bad))
            def make(h, w, scales, bias):
                arith = self.vsl_arith
                if arith == "auto":
                    arith = VF.calibrate_arith(opt.batch_size, h, w, self.device)
                return VF.FusedLossPlan(opt.batch_size, h, w, scales, len(opt.frame_ids) - 1, opt.min_depth,
                                        opt.max_depth, opt.disparity_smoothness, flags=flags, arith=arith,
                                        image_dtype=self.vsl_image_dtype, smooth_level_bias=bias)
            if v1:
                # --v1_multiscale (trainer.py:497-498, 604-605): every level is warped and scored at its own
                # resolution with its own K / images, i.e. S independent single-level problems
                plan = (key, [make(opt.height >> s, opt.width >> s, [0], s) for s in opt.scales])
            else:
                plan = (key, make(opt.height, opt.width, opt.scales, 0))
            self._vsl_plan_cache = plan
        return plan[1]

    def _vsl_level_plans(self, image_dtype=None):
        """[(plan, scale_index_in_plan, source_scale)] per entry of opt.scales."""
        plan = self._vsl_plan(image_dtype)
        if isinstance(plan, list):
            return [(pl, 0, s) for pl, s in zip(plan, self.opt.scales)]
        return [(plan, si, 0) for si, _ in enumerate(self.opt.scales)]

    def _vsl_predictive_masks(self, outputs, scales=None):
        """--predictive_mask (trainer.py:635-647; only used by the reference together with
        --disable_automasking): the mask network's outputs at the warp resolution, and the reference's
        0.2 * BCE(mask, 1) weighting term per scale (plain torch, as in the reference)."""
        opt = self.opt
        if not (getattr(opt, "predictive_mask", False) and getattr(opt, "disable_automasking", False)):
            return None, None
        masks, weighting = [], []
        for scale in (opt.scales if scales is None else scales):
            mask = outputs["predictive_mask"][("disp", scale)]
            if not getattr(opt, "v1_multiscale", False):
                mask = torch.nn.functional.interpolate(mask, [opt.height, opt.width], mode="bilinear",
                                                       align_corners=False)
            masks.append(mask)
            weighting.append((0.2 * torch.nn.functional.binary_cross_entropy(mask, torch.ones_like(mask))).mean())
        return masks, weighting

    def _vsl_poses(self, inputs, outputs, scales=None):
        """T_f per source frame: stereo_T or cam_T_cam (reference trainer.py:510-513).  With posecnn the
        pose depends on the scale (trainer.py:516-525: translation times the level's mean inverse depth) and
        a list per scale is returned; that branch is plain torch on [B,1,H,W] tensors, as in the reference."""
        opt = self.opt
        if getattr(opt, "pose_model_type", "separate_resnet") != "posecnn":
            return [inputs["stereo_T"] if f == "s" else outputs[("cam_T_cam", 0, f)] for f in opt.frame_ids[1:]]
        plan = getattr(self, "_vsl_plan_cache", (None, None))[1]
        if (self.vsl_posecnn_tail == "fused" and scales is None and plan is not None and not isinstance(plan, list)
                and outputs[("disp", opt.scales[0])].is_cuda):
            frames = opt.frame_ids[1:]
            return VF.posecnn_poses(plan, [outputs[("axisangle", 0, f)][:, 0] for f in frames],
                                    [outputs[("translation", 0, f)][:, 0] for f in frames], [f < 0 for f in frames],
                                    [outputs[("disp", s)] for s in opt.scales])
        from .layers import disp_to_depth, transformation_from_parameters
        per_scale = []
        for scale in (opt.scales if scales is None else scales):
            disp = outputs[("disp", scale)]
            if not getattr(opt, "v1_multiscale", False):
                disp = torch.nn.functional.interpolate(disp, [opt.height, opt.width], mode="bilinear", align_corners=False)
            _, depth = disp_to_depth(disp, opt.min_depth, opt.max_depth)
            mean_inv_depth = (1 / depth).mean(3, True).mean(2, True)
            per_scale.append([transformation_from_parameters(
                outputs[("axisangle", 0, f)][:, 0], outputs[("translation", 0, f)][:, 0] * mean_inv_depth[:, 0], f < 0)
                for f in opt.frame_ids[1:]])
        return per_scale

    def _vsl_projections(self, inputs, outputs, source_scale=0):
        """P_f = (K @ T_f)[:, :3, :] per source frame (reference layers.py:254; T as trainer.py:510-513)."""
        K = inputs[("K", source_scale)]
        return [torch.matmul(K, T)[:, :3, :] for T in self._vsl_poses(inputs, outputs)]

    # -- reference surface ----------------------------------------------------------------------
    def generate_images_pred(self, inputs, outputs):
        """Reference trainer.py:491-541.  See the module docstring for ``vsl_side_outputs``."""
        plan = self._vsl_plan(inputs[("color", 0, 0)].dtype)  # validates the options early, like the reference would
        self._vsl_side_pending = None
        mode = self.vsl_side_outputs
        if mode == "fused" and (isinstance(plan, list) or getattr(self.opt, "pose_model_type", "") == "posecnn"):
            mode = "eager"   # per-level plans (--v1_multiscale) and per-scale poses keep the separate kernel
        if mode == "eager":
            self.materialize_side_outputs(inputs, outputs)
        elif mode == "fused":
            self._allocate_fused_side_outputs(inputs, outputs, plan)
        elif mode != "none":
            raise ValueError("vsl_side_outputs must be 'eager', 'fused' or 'none'")

    def _allocate_fused_side_outputs(self, inputs, outputs, plan):
        """``vsl_side_outputs = "fused"``: the tensors are put into ``outputs`` here and FILLED by the fused
        kernel during the following ``compute_losses`` (which forms depth, sampling grid and warped colours
        anyway): the reference-visible outputs at +45 us per step instead of +250 us for separate launches.
        Between the two calls they are uninitialised; the reference never reads them there (trainer.py:399-401)."""
        opt = self.opt
        B, H, W = opt.batch_size, opt.height, opt.width
        dev = inputs[("color", 0, 0)].device
        frames = opt.frame_ids[1:]
        side = {"depth": [], "sample": [], "color": []}
        for scale in opt.scales:
            d = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
            outputs[("depth", 0, scale)] = d
            side["depth"].append(d)
            ss, cc = [], []
            for f in frames:
                smp = torch.empty(B, H, W, 2, dtype=torch.float32, device=dev)
                col = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
                outputs[("sample", f, scale)] = smp
                outputs[("color", f, scale)] = col
                if plan.automask:  # trainer.py:539-541
                    outputs[("color_identity", f, scale)] = inputs[("color", f, 0)]
                ss.append(smp)
                cc.append(col)
            side["sample"].append(ss)
            side["color"].append(cc)
        self._vsl_side_pending = side

    def materialize_side_outputs(self, inputs, outputs):
        """Write ("depth",0,s), ("sample",f,s), ("color",f,s), ("color_identity",f,s) into ``outputs``."""
        with torch.no_grad():
            posecnn = getattr(self.opt, "pose_model_type", "") == "posecnn"
            for (plan, si, src_scale), scale in zip(self._vsl_level_plans(inputs[("color", 0, 0)].dtype), self.opt.scales):
                if posecnn:
                    Ps = [torch.matmul(inputs[("K", src_scale)], T)[:, :3, :]
                          for T in self._vsl_poses(inputs, outputs, [scale])[0]]
                else:
                    Ps = self._vsl_projections(inputs, outputs, src_scale)
                sources = [inputs[("color", f, src_scale)] for f in self.opt.frame_ids[1:]]
                depth, samples, colors = VF.warp_side_outputs(
                    plan, si, outputs[("disp", scale)], inputs[("inv_K", src_scale)], Ps, sources)
                outputs[("depth", 0, scale)] = depth
                for fi, frame_id in enumerate(self.opt.frame_ids[1:]):
                    outputs[("sample", frame_id, scale)] = samples[fi]
                    outputs[("color", frame_id, scale)] = colors[fi]
                    if plan.automask:  # trainer.py:539-541
                        outputs[("color_identity", frame_id, scale)] = inputs[("color", frame_id, src_scale)]

    def compute_reprojection_loss(self, pred, target):
        """Reference trainer.py:543-555: 0.85 * mean_c SSIM + 0.15 * mean_c L1 -> [B,1,H,W]."""
        return VF.reprojection_loss(pred, target, no_ssim=bool(self.opt.no_ssim))

    def compute_losses(self, inputs, outputs):
        """Reference trainer.py:557-686: returns the loss dict, writes ``identity_selection/s``."""
        opt = self.opt
        plan = self._vsl_plan(inputs[("color", 0, 0)].dtype)
        if isinstance(plan, list):
            return self._compute_losses_v1(inputs, outputs, plan)
        S, F = len(opt.scales), len(opt.frame_ids) - 1
        targets = [inputs[("color", 0, s)] for s in opt.scales]
        sources = [inputs[("color", f, 0)] for f in opt.frame_ids[1:]]
        disps = [outputs[("disp", s)] for s in opt.scales]
        dev = disps[0].device
        # one draw per scale, same shape/order/device as trainer.py:656-657
        prefetch = plan.automask and self.vsl_noise_prefetch
        fork = None
        if prefetch:
            # the draws for THIS call were made during the previous one (first call: made here); the point where
            # the next call's draws may start is recorded before the loss kernels are enqueued
            noise = self._vsl_noise_ahead if self._vsl_noise_ahead is not None else self._vsl_draw_noise(plan, dev, S)
            self._vsl_noise_ahead = None
            fork = torch.cuda.Event()
            fork.record(torch.cuda.current_stream(dev))
        else:
            noise = self._vsl_draw_noise(plan, dev, S) if plan.automask else None
        pmasks, weighting = self._vsl_predictive_masks(outputs)
        side = getattr(self, "_vsl_side_pending", None)
        self._vsl_side_pending = None
        vec, masks = VF.fused_loss(plan, targets, sources, disps, inputs[("inv_K", 0)], None, noise,
                                   K=inputs[("K", 0)], Ts=self._vsl_poses(inputs, outputs), predictive_masks=pmasks,
                                   side=side)
        if prefetch:
            # enqueued BEHIND the loss kernels but dependent only on `fork`: the generator kernels fill the SMs the
            # loss kernel's last wave leaves idle instead of running alone in front of the next step
            self._vsl_noise_ahead = self._vsl_draw_noise(plan, dev, S, after=fork, out=self._vsl_noise_out)
        # the whole loss dict as one contiguous device vector (min_loss/s..., loss/s..., loss): a logger can
        # read it back with one copy instead of one per entry
        self.vsl_last_loss_vector = vec.detach() if weighting is None else None
        losses = {}
        for si, scale in enumerate(opt.scales):
            losses["min_loss/{}".format(scale)] = vec[si]
            losses["loss/{}".format(scale)] = vec[S + si] if weighting is None else vec[S + si] + weighting[si]
            if plan.automask:  # the reference writes the mask only with automasking on (trainer.py:668-670)
                outputs["identity_selection/{}".format(scale)] = masks[si]
        losses["loss"] = vec[2 * S] if weighting is None else vec[2 * S] + sum(weighting) / S
        self._vsl_gan_prior(inputs, outputs, losses)
        return losses

    vsl_parallel_noise = True   # the S draws run as parallel branches (side streams); False: back to back on one stream
    # Opt-in software pipelining of the tie-break noise: every compute_losses call consumes the noise drawn during
    # the previous call and draws the next call's while its own loss kernels run.  The k-th call still receives the
    # k-th group of S randn draws of the global generator, in the reference's order (trainer.py:656-657) -- what
    # changes is WHEN the generator is advanced (one call early), so it is off by default: a caller that draws other
    # CUDA random numbers between steps, or checkpoints the generator state, sees a different interleaving.
    vsl_noise_prefetch = False
    _vsl_noise_ahead = None     # noise drawn ahead for the next call (list of S tensors)
    _vsl_noise_out = None       # optional static buffers the ahead draws are written to (graph.GraphedLossStep)

    def _vsl_draw_noise(self, plan, dev, S, after=None, out=None):
        """The tie-break noise of trainer.py:656-657: one ``randn`` per scale, in the reference's order, so the
        global Philox stream is consumed exactly like the reference does (the offsets are assigned on the host, call
        by call).  The S generator kernels are independent; enqueued on side streams they run as parallel branches
        (also inside a captured CUDA graph) instead of four launch-latency-bound kernels in a row."""
        shape = (plan.batch, plan.noise_channels, plan.height, plan.width)
        if after is None and (not self.vsl_parallel_noise or S == 1):
            return [torch.randn(shape, device=dev) for _ in range(S)]
        cur = torch.cuda.current_stream(dev)
        nside = S if after is not None else S - 1   # drawing ahead: every draw leaves the current stream
        side = getattr(self, "_vsl_noise_streams", None)
        if side is None or len(side) < nside or side[0].device != dev:
            side = self._vsl_noise_streams = [torch.cuda.Stream(device=dev) for _ in range(max(nside, 1))]
        noise = out if out is not None else [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(S)]   # allocated on `cur`
        if len(noise) != S or any(t.shape != torch.Size(shape) or t.dtype != torch.float32 for t in noise):
            raise ValueError("noise buffers must be %d float32 tensors of shape %s" % (S, (shape,)))
        fork = after
        if fork is None:
            fork = torch.cuda.Event()
            fork.record(cur)
            noise[0].normal_()                     # randn == empty + normal_(0, 1): same generator calls, same order
        joins = []
        for s in range(0 if after is not None else 1, S):
            st = side[s if after is not None else s - 1]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                noise[s].normal_()
                ev = torch.cuda.Event()
                ev.record(st)
                joins.append(ev)
        for ev in joins:
            cur.wait_event(ev)
        return noise

    def _vsl_gan_prior(self, inputs, outputs, losses):
        """--pre_trained_generator (trainer.py:565-583, :684): the scale-invariant log loss between the generator's
        disparity and every up-sampled ("disp", s), 0.002 / num_scales of their sum added to the total.  The
        generator and its input transform are the caller's (``self.models["pre_trained_generator"]``,
        ``self.gen_transform``, trainer.py:118-131); the loss itself is the SLlog CUDA kernel."""
        opt = self.opt
        if not getattr(opt, "pre_trained_generator", False):
            return
        models = getattr(self, "models", None)
        if not models or "pre_trained_generator" not in models or not hasattr(self, "gen_transform"):
            raise RuntimeError('--pre_trained_generator needs self.models["pre_trained_generator"] and self.gen_transform '
                               "(trainer.py:118-131)")
        from .layers import SLlog, depth_to_disp
        fake_B1 = models["pre_trained_generator"](self.gen_transform(inputs[("color", 0, 0)]))
        _, fake_disp_scaled = depth_to_disp(fake_B1)
        si_loss = getattr(self, "si_loss", None) or SLlog()
        gan_total = 0
        for scale in opt.scales:
            disp = torch.nn.functional.interpolate(outputs[("disp", scale)], [opt.height, opt.width], mode="bilinear",
                                                   align_corners=False)
            gan_loss = si_loss(fake_disp_scaled.contiguous(), disp)
            losses["gan_loss/{}".format(scale)] = gan_loss
            gan_total = gan_total + gan_loss
        losses["loss"] = losses["loss"] + gan_total / len(opt.scales) * 0.002
        self.vsl_last_loss_vector = None   # the dict no longer is one contiguous vector

    depth_metric_names = ["de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3"]  # trainer.py:255-256

    def compute_depth_losses(self, inputs, outputs, losses):
        """Reference trainer.py:688-716 (validation-time depth metrics against inputs["depth_gt"]): up-sample,
        Garg/Eigen crop, median scaling, clamp and the seven metrics in CUDA kernels; one device-to-host read."""
        depth = outputs.get(("depth", 0, 0))
        if depth is None:   # side outputs were skipped ("none" mode): only the depth of scale 0 is needed here
            from .layers import disp_to_depth
            _, depth = disp_to_depth(outputs[("disp", 0)].detach(), self.opt.min_depth, self.opt.max_depth)
        vals = VF.depth_losses(depth, inputs["depth_gt"]).cpu().numpy()
        for i, metric in enumerate(self.depth_metric_names):
            losses[metric] = np.array(vals[i])


    def _compute_losses_v1(self, inputs, outputs, plans):
        """--v1_multiscale: one fused launch per level at that level's resolution (trainer.py:604-605)."""
        opt = self.opt
        F = len(opt.frame_ids) - 1
        losses, total = {}, 0
        for plan, scale in zip(plans, opt.scales):
            disp = outputs[("disp", scale)]
            noise = None
            if plan.automask:   # one level per call: a single draw
                noise = [torch.randn((opt.batch_size, plan.noise_channels, plan.height, plan.width), device=disp.device)]
            pmasks, weighting = self._vsl_predictive_masks(outputs, [scale])
            vec, masks = VF.fused_loss(plan, [inputs[("color", 0, scale)]],
                                       [inputs[("color", f, scale)] for f in opt.frame_ids[1:]], [disp],
                                       inputs[("inv_K", scale)], None, noise, K=inputs[("K", scale)],
                                       Ts=self._vsl_poses(inputs, outputs, [scale]), predictive_masks=pmasks)
            level_loss = vec[1] if weighting is None else vec[1] + weighting[0]
            losses["min_loss/{}".format(scale)] = vec[0]
            losses["loss/{}".format(scale)] = level_loss
            if plan.automask:
                outputs["identity_selection/{}".format(scale)] = masks[0]
            total = total + level_loss
        losses["loss"] = total / self.num_scales if hasattr(self, "num_scales") else total / len(opt.scales)
        self._vsl_gan_prior(inputs, outputs, losses)
        return losses


class LossPath(ViewSynthesisLossMixin):
    """Stand-alone carrier of the path: ``LossPath(opt).generate_images_pred / compute_losses``.

    ``opt`` needs the fields the reference methods read: height, width, batch_size, scales, frame_ids,
    min_depth, max_depth, disparity_smoothness and the ablation flags (options.py:59-159).
    """

    def __init__(self, opt, device="cuda", side_outputs="eager", arith="auto"):
        self.opt = opt
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.VslError("LossPath runs on CUDA only (the reference's --no_cuda path has no "
                                "counterpart here; use the reference itself for CPU)")
        self.num_scales = len(opt.scales)
        self.vsl_side_outputs = side_outputs
        self.vsl_arith = arith
        _lib.load()


def make_opt(**kw):
    """Namespace with the reference's defaults for the fields the path reads (options.py)."""
    opt = types.SimpleNamespace(
        height=192, width=640, batch_size=12, scales=[0, 1, 2, 3], frame_ids=[0, -1, 1],
        min_depth=0.1, max_depth=150.0, disparity_smoothness=1e-4,
        v1_multiscale=False, avg_reprojection=False, disable_automasking=False, predictive_mask=False,
        no_ssim=False, pose_model_type="separate_resnet", pre_trained_generator=False, no_cuda=False)
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt
