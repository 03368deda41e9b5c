"""torch.autograd wrappers over the C ABI (include/vsl.h).  PyTorch is used for device memory,
streams and autograd plumbing only; every arithmetic step runs in libvsl_b200.so.

The functions refuse CPU tensors: there is no CPU implementation of this path in the product.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import VSL_MAX_SCALES, VSL_MAX_SRC, VslDesc, VslLossBuffers, check, ptr


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise _lib.VslError(
            "%s is on %s: the view-synthesis loss path only runs on CUDA (no CPU fallback)" % (name, t.device))
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


# --------------------------------------------------------------------------------------------------
# rounding-order calibration
# --------------------------------------------------------------------------------------------------
_ARITH_CACHE = {}


def calibrate_arith(batch, height, width, device):
    """Which accumulation order does torch.bmm (cuBLAS) use for this shape on this device?

    The reference computes rays and projections with ``torch.matmul`` (layers.py:235, :256).  cuBLAS
    accumulates one FMA chain for batch >= 2, but for batch 1 it may select a kernel that adds un-fused
    products.  Bit-exact projection indices need the same order, so once per (batch, H, W, device) the
    two bmm shapes are run through torch and through ``vsl_probe_bmm`` variants; the matching
    VSL_ARITH_* bits are returned.  Runs at plan construction only, never on the hot path."""
    device = torch.device(device)
    key = (batch, height, width, device.index if device.index is not None else torch.cuda.current_device())
    if key in _ARITH_CACHE:
        return _ARITH_CACHE[key]
    lib = _lib.load()
    n = height * width
    gen = torch.Generator(device="cpu").manual_seed(1234)
    M = torch.randn(batch, 4, 4, generator=gen).to(device)
    bits = 0
    # (k, columns, probe variants -> plan bits): rays [B,3,3]x[B,3,HW]; projection [B,3,4]x[B,4,HW];
    # P = (K@T)[:, :3, :], a [B,4,4]x[B,4,4] product (probed through its first three rows)
    probes = ((3, n, False, {0: 0, _lib.ARITH_DOT3_NOFMA: _lib.ARITH_DOT3_NOFMA, _lib.ARITH_DOT3_REVERSE: _lib.ARITH_DOT3_REVERSE}),
              (4, n, False, {0: 0, _lib.ARITH_DOT_NOFMA: _lib.ARITH_DOT_NOFMA, _lib.ARITH_DOT_REVERSE: _lib.ARITH_DOT_REVERSE}),
              (4, 4, True, {0: 0, _lib.ARITH_DOT_NOFMA: _lib.ARITH_DOTKT_NOFMA, _lib.ARITH_DOT_REVERSE: _lib.ARITH_DOTKT_REVERSE}))
    for k, n, full_kt, variants in probes:   # full_kt: the [B,4,4]x[B,4,4] product (a 2x2 image also has n == 4)
        X = torch.randn(batch, k, n, generator=gen).to(device)
        A = M[:, :3, :k]                       # sliced like inv_K[:, :3, :3] / (K@T)[:, :3, :]
        tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            ref = torch.matmul(M, X)[:, :3, :].contiguous() if full_kt else torch.matmul(A, X)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32
        Ac = A.contiguous()
        out = torch.empty_like(ref)
        chosen = None
        for v, plan_bits in variants.items():
            check(lib.vsl_probe_bmm(batch, k, n, v, Ac.data_ptr(), X.data_ptr(), out.data_ptr(), _stream()),
                  "vsl_probe_bmm")
            if torch.equal(out, ref):
                chosen = plan_bits
                break
        if chosen is None:
            raise _lib.VslError("could not reproduce torch.bmm's rounding for [%d,3,%d]x[%d,%d,%d] on this device; "
                                "pass an explicit arith" % (batch, k, batch, k, n))
        bits |= chosen
    _ARITH_CACHE[key] = bits
    return bits


_POSE_ARITH_CACHE = {}


def calibrate_pose_arith(batch, device):
    """Rounding order of the reference's pose -> 4x4 ops for this batch size on this device: the
    [B,4,4]x[B,4,4] bmm (shared with K@T) and torch.norm over the three axis-angle components (a 4-lane
    shuffle tree on CUDA).  Checked once per (batch, device) against torch; raises if neither known order
    reproduces it."""
    device = torch.device(device)
    key = (batch, device.index if device.index is not None else torch.cuda.current_device())
    if key in _POSE_ARITH_CACHE:
        return _POSE_ARITH_CACHE[key]
    lib = _lib.load()
    gen = torch.Generator().manual_seed(7)
    # the only inexact product of the conversion is matmul(R.transpose(1, 2), T(-t)) of the inverted pose
    # (layers.py:104-111): a TRANSPOSED left operand, for which cuBLAS may pick another kernel than for K@T
    Rm, Tm = torch.randn(batch, 4, 4, generator=gen).to(device), torch.randn(batch, 4, 4, generator=gen).to(device)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = torch.matmul(Rm.transpose(1, 2), Tm)[:, :3, :].contiguous()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    A = Rm.transpose(1, 2)[:, :3, :].contiguous()
    out = torch.empty_like(ref)
    bits = None
    for probe, plan_bits in ((0, 0), (_lib.ARITH_DOT_NOFMA, _lib.ARITH_DOTKT_NOFMA), (_lib.ARITH_DOT_REVERSE, _lib.ARITH_DOTKT_REVERSE)):
        check(lib.vsl_probe_bmm(batch, 4, 4, probe, A.data_ptr(), Tm.data_ptr(), out.data_ptr(), _stream()), "vsl_probe_bmm")
        if torch.equal(out, ref):
            bits = plan_bits
            break
    if bits is None:
        raise _lib.VslError("could not reproduce torch.matmul(R^T, T) rounding for batch %d on this device" % batch)
    v = (0.3 * torch.randn(batch, 1, 3, generator=gen)).to(device)
    ref = torch.norm(v, 2, 2, True).view(-1)
    sq = (v * v).view(-1, 3)
    tree, seq = torch.sqrt((sq[:, 0] + sq[:, 2]) + sq[:, 1]), torch.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2])
    if torch.equal(tree, ref):
        pass
    elif torch.equal(seq, ref):
        bits |= _lib.ARITH_NORM_SEQ
    else:
        raise _lib.VslError("could not reproduce torch.norm's rounding for [%d,1,3] on this device" % batch)
    _POSE_ARITH_CACHE[key] = bits
    return bits


class _PoseMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axisangle, translation, invert, arith):
        aa, tr = _dev(axisangle, "axisangle"), _dev(translation, "translation")
        B = aa.shape[0]
        if aa.numel() != 3 * B or tr.numel() != 3 * B:
            raise ValueError("axisangle / translation must be [B,1,3], got %s / %s" % (tuple(aa.shape), tuple(tr.shape)))
        T = torch.empty(B, 4, 4, dtype=torch.float32, device=aa.device)
        check(_lib.load().vsl_pose_forward(B, int(invert), arith, aa.data_ptr(), tr.data_ptr(), T.data_ptr(), _stream()),
              "vsl_pose_forward")
        ctx.save_for_backward(aa, tr)
        ctx.invert = int(invert)
        return T

    @staticmethod
    def backward(ctx, g):
        aa, tr = ctx.saved_tensors
        g = _dev(g, "grad")
        gaa, gtr = torch.empty_like(aa), torch.empty_like(tr)
        check(_lib.load().vsl_pose_backward(aa.shape[0], ctx.invert, aa.data_ptr(), tr.data_ptr(), g.data_ptr(),
                                            gaa.data_ptr(), gtr.data_ptr(), _stream()), "vsl_pose_backward")
        return gaa, gtr, None, None


def pose_matrix(axisangle, translation, invert=False, arith="auto"):
    """transformation_from_parameters (layers.py:97-114) in one kernel, bit-identical to the torch ops."""
    if arith == "auto":
        arith = calibrate_pose_arith(axisangle.shape[0], axisangle.device)
    return _PoseMatrix.apply(axisangle, translation, bool(invert), arith)


class _PoseCnnTail(torch.autograd.Function):
    """trainer.py:516-525 for every scale and frame: T[s][f] from (axisangle_f, translation_f, disp_s)."""

    @staticmethod
    def forward(ctx, plan, inverts, arith, n_frames, *tensors):
        S, F, B = len(plan.scales), n_frames, plan.batch
        aa = [_dev(t, "axisangle") for t in tensors[:F]]
        tr = [_dev(t, "translation") for t in tensors[F:2 * F]]
        disps = [_dev(t, "disp") for t in tensors[2 * F:]]
        for t in aa + tr:
            if t.numel() != 3 * B:
                raise ValueError("axisangle / translation must hold [B,3] values, got %s" % (tuple(t.shape),))
        for s, d in enumerate(disps):
            if tuple(d.shape) != plan.level_shapes[s]:
                raise ValueError("disp[%d] has shape %s, plan expects %s" % (s, tuple(d.shape), plan.level_shapes[s]))
        dev = disps[0].device
        lib = plan.lib
        T = torch.empty(S, F, B, 4, 4, dtype=torch.float32, device=dev)
        mean_inv = torch.empty(S, B, dtype=torch.float32, device=dev)
        nbytes = lib.vsl_posecnn_workspace_bytes(ctypes.byref(plan.desc))
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
        dp, ap, tp = (ctypes.c_void_p * VSL_MAX_SCALES)(), (ctypes.c_void_p * VSL_MAX_SRC)(), (ctypes.c_void_p * VSL_MAX_SRC)()
        inv = (ctypes.c_int32 * VSL_MAX_SRC)()
        for s in range(S):
            dp[s] = disps[s].data_ptr()
        for f in range(F):
            ap[f], tp[f], inv[f] = aa[f].data_ptr(), tr[f].data_ptr(), int(inverts[f])
        check(lib.vsl_posecnn_forward(ctypes.byref(plan.desc), ctypes.byref(dp), F, ctypes.byref(ap), ctypes.byref(tp),
                                      ctypes.byref(inv), int(arith), T.data_ptr(), mean_inv.data_ptr(), ws.data_ptr(),
                                      nbytes, _stream()), "vsl_posecnn_forward")
        ctx.save_for_backward(mean_inv, *aa, *tr)
        ctx.meta = (plan, list(inverts), F, [t.shape for t in tensors[:2 * F]])
        return T

    @staticmethod
    def backward(ctx, gT):
        plan, inverts, F, shapes = ctx.meta
        S, B = len(plan.scales), plan.batch
        mean_inv = ctx.saved_tensors[0]
        aa, tr = ctx.saved_tensors[1:1 + F], ctx.saved_tensors[1 + F:1 + 2 * F]
        gT = _dev(gT, "grad_T")
        dev = gT.device
        gaa = [torch.empty(B, 3, dtype=torch.float32, device=dev) for _ in range(F)]
        gtr = [torch.empty(B, 3, dtype=torch.float32, device=dev) for _ in range(F)]
        gconst = torch.empty(S, B, dtype=torch.float32, device=dev)
        ap, tp = (ctypes.c_void_p * VSL_MAX_SRC)(), (ctypes.c_void_p * VSL_MAX_SRC)()
        gap, gtp = (ctypes.c_void_p * VSL_MAX_SRC)(), (ctypes.c_void_p * VSL_MAX_SRC)()
        inv = (ctypes.c_int32 * VSL_MAX_SRC)()
        for f in range(F):
            ap[f], tp[f], inv[f] = aa[f].data_ptr(), tr[f].data_ptr(), int(inverts[f])
            gap[f], gtp[f] = gaa[f].data_ptr(), gtr[f].data_ptr()
        check(plan.lib.vsl_posecnn_backward(ctypes.byref(plan.desc), F, ctypes.byref(ap), ctypes.byref(tp), ctypes.byref(inv),
                                            mean_inv.data_ptr(), gT.data_ptr(), ctypes.byref(gap), ctypes.byref(gtp),
                                            gconst.data_ptr(), _stream()), "vsl_posecnn_backward")
        gd = [gconst[s].view(B, 1, 1, 1).expand(plan.level_shapes[s]) for s in range(S)]
        grads = [g.view(sh) for g, sh in zip(gaa + gtr, shapes)]
        return (None, None, None, None) + tuple(grads) + tuple(gd)


def posecnn_poses(plan, axisangles, translations, inverts, disps, arith="auto"):
    """[[T_{s,f} for f] for s]: the posecnn pose tail (trainer.py:516-525) in two launches (+ one backward launch)."""
    if arith == "auto":
        arith = calibrate_pose_arith(plan.batch, disps[0].device)
    F = len(axisangles)
    T = _PoseCnnTail.apply(plan, list(inverts), arith, F, *(list(axisangles) + list(translations) + list(disps)))
    return [[T[s, f] for f in range(F)] for s in range(len(plan.scales))]


# --------------------------------------------------------------------------------------------------
# fused loss
# --------------------------------------------------------------------------------------------------
class FusedLossPlan:
    """What Trainer.__init__ fixes once for the path (reference trainer.py:245-259): sizes, scales,
    number of source frames, depth range, smoothness weight.  Owns the workspace; one call at a time
    per plan (calls are stream-ordered)."""

    def __init__(self, batch, height, width, scales, num_src, min_depth, max_depth,
                 disparity_smoothness, flags=_lib.FLAG_AUTOMASK, arith=0, image_dtype=torch.float32,
                 smooth_level_bias=0):
        scales = list(scales)
        if not (1 <= len(scales) <= VSL_MAX_SCALES):
            raise ValueError("1..%d scales supported" % VSL_MAX_SCALES)
        if not (1 <= num_src <= VSL_MAX_SRC):
            raise ValueError("1..%d source frames supported" % VSL_MAX_SRC)
        self.batch, self.height, self.width = int(batch), int(height), int(width)
        self.scales, self.num_src = scales, int(num_src)
        d = VslDesc()
        d.abi_version = _lib.VSL_ABI_VERSION
        d.batch, d.height, d.width = self.batch, self.height, self.width
        d.num_scales = len(scales)
        for i, s in enumerate(scales):
            d.scale_ids[i] = int(s)
        d.num_src = self.num_src
        d.flags = int(flags)
        if image_dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("colour images must be stored as float32 or bfloat16")
        self.image_dtype = image_dtype
        self.automask = bool(flags & _lib.FLAG_AUTOMASK)
        # channels of the tie-break noise: the identity losses are averaged over frames first with
        # --avg_reprojection (trainer.py:629-630), so the reference draws [B,1,H,W] there
        self.noise_channels = 1 if (flags & _lib.FLAG_AVG_REPROJECTION) else self.num_src
        d.image_dtype = _lib.DTYPE_BF16 if image_dtype == torch.bfloat16 else _lib.DTYPE_F32
        d.arith = int(arith)
        # Python-double scalars rounded to fp32 at the op, as PyTorch does (layers.py:90-93)
        d.min_disp = float(np.float32(1.0 / max_depth))
        d.disp_range = float(np.float32(1.0 / min_depth - 1.0 / max_depth))
        d.eps = float(np.float32(1e-7))
        d.smooth_weight = float(disparity_smoothness)
        d.smooth_level_bias = int(smooth_level_bias)
        self.desc = d
        self.lib = _lib.load()
        self.ws_bytes = self.lib.vsl_loss_workspace_bytes(ctypes.byref(d))
        if self.ws_bytes == 0:
            raise _lib.VslError("invalid problem descriptor (sizes must be divisible by 2**scale)")
        self._ws = {}
        self.kernel_events = None  # set to a KernelEvents() to time the photometric kernel per call
        self.level_shapes = [(self.batch, 1, self.height >> s, self.width >> s) for s in scales]

    def workspace(self, device):
        ws = self._ws.get(device)
        if ws is None:
            ws = torch.empty(self.ws_bytes // 4, dtype=torch.float32, device=device)
            check(self.lib.vsl_loss_workspace_init(ctypes.byref(self.desc), ws.data_ptr(), self.ws_bytes, _stream()),
                  "vsl_loss_workspace_init")
            self._ws[device] = ws
        return ws


class KernelEvents:
    """CUDA event pairs recorded by the library around the photometric kernel (bench instrumentation)."""

    def __init__(self):
        self.lib = _lib.load()
        self.pairs = []

    def new_pair(self):
        a, b = ctypes.c_void_p(), ctypes.c_void_p()
        check(self.lib.vsl_event_create(ctypes.byref(a)), "vsl_event_create")
        check(self.lib.vsl_event_create(ctypes.byref(b)), "vsl_event_create")
        self.pairs.append((a, b))
        return a, b

    def drain_ms(self):
        """Elapsed ms of every recorded pair (waits for them); the pairs are destroyed."""
        out = []
        for a, b in self.pairs:
            ms = ctypes.c_float()
            check(self.lib.vsl_event_elapsed_ms(a, b, ctypes.byref(ms)), "vsl_event_elapsed_ms")
            out.append(ms.value)
            self.lib.vsl_event_destroy(a)
            self.lib.vsl_event_destroy(b)
        self.pairs = []
        return out


def _side_ptr(t, shape, dev):
    if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
        raise _lib.VslError("side output buffer must be a contiguous float32 %s tensor on %s (got %s %s %s)"
                            % (tuple(shape), dev, t.dtype, tuple(t.shape), t.device))
    return t.data_ptr()


class _FusedLoss(torch.autograd.Function):
    """losses vector [2S+1] = (min_loss/s ..., loss/s ..., loss), masks...  <- disps, P matrices."""

    @staticmethod
    def forward(ctx, plan, targets, sources, inv_K, noise, want_mask, K, use_T, n_pmask, side, fwd_only, n_srcgrad,
                *leaves):
        S, F = len(plan.scales), plan.num_src
        # n_srcgrad == F: the last F leaves are the source images themselves (they require grad: the optional
        # source-image gradient, see backward); they are read through `sources` like always
        if n_srcgrad:
            leaves = leaves[:len(leaves) - n_srcgrad]
        # the masks are non-differentiable outputs: without this autograd hands backward() a zero-filled
        # [B,H,W] tensor per mask (four 5.9 MB fill kernels per step at config 1)
        ctx.set_materialize_grads(False)
        # fwd_only (decided by the caller: grad mode is always off inside forward()): nothing to differentiate —
        # Trainer.val() under torch.no_grad(), trainer.py:463-489 — so the kernel keeps no adjoint state and
        # writes no gradient
        pmasks = leaves[len(leaves) - n_pmask:] if n_pmask else ()   # --predictive_mask, one [B,F,H,W] per scale
        leaves = leaves[:len(leaves) - n_pmask] if n_pmask else leaves
        disps, Ps = leaves[:S], leaves[S:]   # Ps: projection matrices [B,3,4], or poses T [B,4,4] if use_T
        per_scale = use_T == "per_scale"     # poses given per (scale, frame): S*F tensors, scale-major
        dev = disps[0].device
        B, H, W = plan.batch, plan.height, plan.width
        buf = VslLossBuffers()
        keep, src_c, tgt0_c = [], [], None
        for s in range(S):
            t = _dev(targets[s], "target[%d]" % s, plan.image_dtype)
            d = _dev(disps[s], "disp[%d]" % s)
            if tuple(d.shape) != plan.level_shapes[s]:
                raise ValueError("disp[%d] has shape %s, plan expects %s" % (s, tuple(d.shape), plan.level_shapes[s]))
            if tuple(t.shape) != (B, 3, H >> plan.scales[s], W >> plan.scales[s]):
                raise ValueError("target[%d] has shape %s" % (s, tuple(t.shape)))
            buf.target[s], buf.disp[s] = t.data_ptr(), d.data_ptr()
            keep += [t, d]
            if s == 0:
                tgt0_c = t
            if plan.automask:
                z = _dev(noise[s], "noise[%d]" % s)
                if tuple(z.shape) != (B, plan.noise_channels, H, W):
                    raise ValueError("noise[%d] has shape %s, expected %s" % (s, tuple(z.shape), (B, plan.noise_channels, H, W)))
                buf.noise[s] = z.data_ptr()
                keep.append(z)
        for f in range(F):
            src = _dev(sources[f], "source[%d]" % f, plan.image_dtype)
            if tuple(src.shape) != (B, 3, H, W):
                raise ValueError("source[%d] has shape %s" % (f, tuple(src.shape)))
            buf.source[f] = src.data_ptr()
            keep.append(src)
            src_c.append(src)
            for s in range(S if per_scale else 1):
                P = _dev(Ps[s * F + f], "T[%d]" % f if use_T else "P[%d]" % f)
                if tuple(P.shape) != ((B, 4, 4) if use_T else (B, 3, 4)):
                    raise ValueError("pose[%d] has shape %s" % (f, tuple(P.shape)))
                if per_scale:
                    buf.T_scale[s][f] = P.data_ptr()
                elif use_T:
                    buf.T[f] = P.data_ptr()
                else:
                    buf.P[f] = P.data_ptr()
                keep.append(P)
        if use_T:
            Kc = _dev(K, "K")
            if tuple(Kc.shape) != (B, 4, 4):
                raise ValueError("K has shape %s" % (tuple(Kc.shape),))
            buf.K = Kc.data_ptr()
            keep.append(Kc)
        iK = _dev(inv_K, "inv_K")
        if tuple(iK.shape) != (B, 4, 4):
            raise ValueError("inv_K has shape %s" % (tuple(iK.shape),))
        buf.inv_K = iK.data_ptr()
        keep.append(iK)

        if n_pmask:
            if plan.automask or n_pmask != S:
                raise ValueError("predictive masks need --disable_automasking and one mask per scale")
            gpm = []
            for s in range(S):
                m = _dev(pmasks[s], "predictive_mask[%d]" % s)
                if tuple(m.shape) != (B, F, H, W):
                    raise ValueError("predictive_mask[%d] has shape %s, expected %s" % (s, tuple(m.shape), (B, F, H, W)))
                buf.predictive_mask[s] = m.data_ptr()
                keep.append(m)
                if not fwd_only:
                    gm = torch.empty_like(m)
                    buf.grad_predictive_mask[s] = gm.data_ptr()
                    gpm.append(gm)
            ctx.gpm = gpm
        ctx.n_pmask = n_pmask
        ctx.n_srcgrad = n_srcgrad
        if n_srcgrad and not fwd_only:
            # the backward needs every warped image and sampling grid plus the arg-min channel per pixel
            if n_pmask:
                raise NotImplementedError("source-image gradients are not implemented together with --predictive_mask")
            if plan.image_dtype != torch.float32:
                raise NotImplementedError("source-image gradients need fp32 image storage")
            side = dict(side or {})
            for name, shape in (("sample", (B, H, W, 2)), ("color", (B, 3, H, W))):
                cur = side.get(name) or [[None] * F for _ in range(S)]
                side[name] = [[cur[s][f] if cur[s][f] is not None else torch.empty(shape, dtype=torch.float32, device=dev)
                               for f in range(F)] for s in range(S)]
            ctx.winner = [torch.empty(B, H, W, dtype=torch.uint8, device=dev) for _ in range(S)]
            for s in range(S):
                buf.winner[s] = ctx.winner[s].data_ptr()
            ctx.src_side = side
            ctx.src_images = (tgt0_c, src_c)
        if side is not None:
            # side outputs of generate_images_pred, written in place by the kernel (no autograd edge: the
            # reference's own use of them is inside the loss this call computes)
            for s in range(S):
                d = side["depth"][s] if side.get("depth") else None
                if d is not None:
                    buf.side_depth[s] = _side_ptr(d, (B, 1, H, W), dev)
                for f in range(F):
                    for name, field, shape in (("sample", buf.side_sample, (B, H, W, 2)), ("color", buf.side_color, (B, 3, H, W))):
                        t = side[name][s][f] if side.get(name) else None
                        if t is not None:
                            field[s][f] = _side_ptr(t, shape, dev)
        # one flat allocation for everything the backward keeps
        n_levels = [B * (H >> s) * (W >> s) for s in plan.scales]
        sizes = [3 * S + 1] if fwd_only else [3 * S + 1, S * F * B * 12, S * B * 2] + n_levels + n_levels
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        parts = torch.split(flat, sizes)
        losses = parts[0]
        buf.losses = losses.data_ptr()
        if not fwd_only:
            gradP, norm = parts[1], parts[2]
            gphoto, gsmooth = parts[3:3 + S], parts[3 + S:3 + 2 * S]
            buf.grad_P, buf.smooth_norm = gradP.data_ptr(), norm.data_ptr()
        masks = []
        for s in range(S):
            if not fwd_only:
                buf.grad_disp_photo[s] = gphoto[s].data_ptr()
                buf.grad_disp_smooth[s] = gsmooth[s].data_ptr()
            if want_mask and plan.automask:
                m = torch.empty(B, H, W, dtype=torch.float32, device=dev)
                buf.mask[s] = m.data_ptr()
                masks.append(m)
        ws = plan.workspace(dev)
        ev = plan.kernel_events.new_pair() if plan.kernel_events is not None else (None, None)
        desc = plan.desc
        if fwd_only:
            desc = type(plan.desc).from_buffer_copy(plan.desc)
            desc.flags |= _lib.FLAG_FORWARD_ONLY
        check(plan.lib.vsl_loss_forward_backward_timed(ctypes.byref(desc), ctypes.byref(buf), ws.data_ptr(),
                                                       plan.ws_bytes, _stream(), ev[0], ev[1]),
              "vsl_loss_forward_backward")
        ctx.plan, ctx.buf, ctx.keep, ctx.use_T = plan, buf, keep, use_T
        # what backward() reads through the raw pointers in `buf`: the unit gradients (flat) and K (dL/dT = K^T dL/dP).
        # Saved through autograd so an in-place overwrite between forward and backward (a stager slot or a graph
        # input re-filled too early) raises instead of silently producing a wrong dL/dT.
        ctx.save_for_backward(flat, Kc if use_T else None)
        ctx.mark_non_differentiable(*masks)
        out = losses[:2 * S + 1]
        ctx.smooth_terms = losses[2 * S + 1:]
        return (out,) + tuple(masks)

    @staticmethod
    def backward(ctx, gvec, *_gmasks):
        plan = ctx.plan
        S, F, B = len(plan.scales), plan.num_src, plan.batch
        if gvec is None:  # no loss entry was used (grads are not materialised, see forward)
            return (None,) * (12 + S + (S if ctx.use_T == "per_scale" else 1) * F + ctx.n_pmask + ctx.n_srcgrad)
        ctx.saved_tensors  # version-counter check of the buffers `ctx.buf` points at
        dev = gvec.device
        up = _dev(gvec, "upstream gradient")
        n_levels = [int(np.prod(sh)) for sh in plan.level_shapes]
        pe = 16 if ctx.use_T else 12   # gradient w.r.t. T [B,4,4] or P [B,3,4]
        n_pose = (S if ctx.use_T == "per_scale" else 1) * F
        flat = torch.empty(sum(n_levels) + n_pose * B * pe, dtype=torch.float32, device=dev)
        parts = torch.split(flat, n_levels + [n_pose * B * pe])
        out_ptrs = (ctypes.c_void_p * VSL_MAX_SCALES)()
        for s in range(S):
            out_ptrs[s] = parts[s].data_ptr()
        gP = parts[S]
        check(plan.lib.vsl_loss_combine_grads(ctypes.byref(plan.desc), up.data_ptr(), ctypes.byref(ctx.buf),
                                              ctypes.byref(out_ptrs), None if ctx.use_T else gP.data_ptr(),
                                              gP.data_ptr() if ctx.use_T else None, _stream()),
              "vsl_loss_combine_grads")
        gd = [parts[s].view(plan.level_shapes[s]) for s in range(S)]
        gPs = [gP.view(n_pose, B, 4, 4)[i] if ctx.use_T else gP.view(F, B, 3, 4)[i] for i in range(n_pose)]
        gpm = ()
        if ctx.n_pmask:
            # d L / d mask_s = a_s * d(min_loss/s)/d mask_s with a_s the upstream weight of min_loss/s
            # (the same a_s vsl_loss_combine_grads uses; tiny torch arithmetic on the [2S+1] vector)
            a = up[:S] + up[S:2 * S] + up[2 * S] / S
            gpm = tuple(ctx.gpm[s] * a[s] for s in range(S))
        gsrc = ()
        if ctx.n_srcgrad:
            gsrc = _source_image_grads(ctx, plan, up)
        return (None,) * 12 + tuple(gd) + tuple(gPs) + gpm + gsrc


def _source_image_grads(ctx, plan, up):
    """d L / d inputs[("color", f, 0)] per source frame (what the reference's autograd returns when the source
    images require grad): identity candidates directly (trainer.py:620-633), warped candidates through the
    bilinear scatter (grid_sample's backward, trainer.py:534-537).  Per frame and scale: what the candidate's
    reprojection loss receives from the loss dict (vsl_source_grad_upstream, from the arg-min channel the fused
    kernel recorded) -> d/d pred (vsl_reprojection_loss_backward) -> scatter (vsl_grid_sample_backward_source)."""
    lib = plan.lib
    S, F, B, H, W = len(plan.scales), plan.num_src, plan.batch, plan.height, plan.width
    dev = up.device
    target, sources = ctx.src_images
    automask = plan.automask
    no_ssim = 1 if (plan.desc.flags & _lib.FLAG_NO_SSIM) else 0
    up_id = torch.empty(F, B, H, W, dtype=torch.float32, device=dev) if automask else None
    up_w = [torch.empty(F, B, H, W, dtype=torch.float32, device=dev) for _ in range(S)]
    wptr = (ctypes.c_void_p * VSL_MAX_SCALES)()
    uptr = (ctypes.c_void_p * VSL_MAX_SCALES)()
    for s in range(S):
        wptr[s], uptr[s] = ctx.winner[s].data_ptr(), up_w[s].data_ptr()
    check(lib.vsl_source_grad_upstream(ctypes.byref(plan.desc), up.data_ptr(), ctypes.byref(wptr), ptr(up_id),
                                       ctypes.byref(uptr), _stream()), "vsl_source_grad_upstream")
    nbytes = 0 if no_ssim else lib.vsl_ssim_workspace_bytes(B, 3, H, W)
    ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
    tmp = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
    out = []
    for f in range(F):
        if not ctx.needs_input_grad[len(ctx.needs_input_grad) - F + f]:
            out.append(None)
            continue
        if automask:
            g = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
            check(lib.vsl_reprojection_loss_backward(B, H, W, no_ssim, sources[f].data_ptr(), target.data_ptr(),
                                                     up_id[f].data_ptr(), g.data_ptr(), None, ws.data_ptr(), nbytes,
                                                     _stream()), "vsl_reprojection_loss_backward")
        else:
            g = torch.zeros(B, 3, H, W, dtype=torch.float32, device=dev)
        for s in range(S):
            color, grid = ctx.src_side["color"][s][f], ctx.src_side["sample"][s][f]
            check(lib.vsl_reprojection_loss_backward(B, H, W, no_ssim, color.data_ptr(), target.data_ptr(),
                                                     up_w[s][f].data_ptr(), tmp.data_ptr(), None, ws.data_ptr(), nbytes,
                                                     _stream()), "vsl_reprojection_loss_backward")
            check(lib.vsl_grid_sample_backward_source(B, H, W, grid.data_ptr(), tmp.data_ptr(), g.data_ptr(), _stream()),
                  "vsl_grid_sample_backward_source")
        out.append(g)
    return tuple(out)


def fused_loss(plan, targets, sources, disps, inv_K, Ps, noise, want_mask=True, K=None, Ts=None,
               predictive_masks=None, side=None):
    """Run the fused path.  Returns (loss_vector[2S+1], [mask_s ...]).

    loss_vector order: min_loss/s for every scale, loss/s for every scale, loss
    (reference trainer.py:672-685).  The camera of each source frame is given either as ``Ps``
    (projection matrices (K@T)[:, :3, :]) or, preferred, as ``K`` + ``Ts`` (the 4x4 poses): the kernel then
    forms K@T itself and the backward returns dL/dT directly, with no torch matmul in between.
    Gradients flow to ``disps`` and to ``Ps`` / ``Ts``.  ``side``: optional pre-allocated float32 buffers
    ``{"depth": [S x [B,1,H,W]], "sample": [S][F x [B,H,W,2]], "color": [S][F x [B,3,H,W]]}`` (entries may be
    None) that receive the reference's ``generate_images_pred`` outputs from the same kernel.
    """
    use_T = Ts is not None
    poses = list(Ts if use_T else Ps)
    if use_T and poses and isinstance(poses[0], (list, tuple)):
        # one pose per (scale, frame) — posecnn (trainer.py:516-525): Ts[s][f]
        use_T = "per_scale"
        poses = [T for per_frame in poses for T in per_frame]
    pm = list(predictive_masks or [])
    leaves = list(disps) + poses + pm
    # optional: gradients with respect to the source images (the reference's autograd gives them whenever
    # inputs[("color", f, 0)] requires grad); the images then travel as autograd leaves as well
    n_srcgrad = len(sources) if torch.is_grad_enabled() and any(t.requires_grad for t in sources) else 0
    if n_srcgrad:
        leaves = leaves + list(sources)
    fwd_only = not (torch.is_grad_enabled() and any(t.requires_grad for t in leaves))
    res = _FusedLoss.apply(plan, list(targets), [t.detach() for t in sources], inv_K, list(noise or []), bool(want_mask),
                           K, use_T, len(pm), side, fwd_only, n_srcgrad, *leaves)
    return res[0], list(res[1:])


def warp_side_outputs(plan, scale_index, disp, inv_K, Ps, sources, want_depth=True, want_sample=True,
                      want_color=True):
    """outputs[("depth",0,s)], [("sample",f,s)], [("color",f,s)] of trainer.py:500-537 for one scale
    (non-differentiable: in the reference they feed the loss, here the fused kernel recomputes them)."""
    B, H, W, F = plan.batch, plan.height, plan.width, plan.num_src
    disp = _dev(disp.detach(), "disp")
    dev = disp.device
    iK = _dev(inv_K, "inv_K")
    Pp = (ctypes.c_void_p * VSL_MAX_SRC)()
    Sp = (ctypes.c_void_p * VSL_MAX_SRC)()
    samp = (ctypes.c_void_p * VSL_MAX_SRC)()
    colp = (ctypes.c_void_p * VSL_MAX_SRC)()
    keep, samples, colors = [], [], []
    for f in range(F):
        P = _dev(Ps[f].detach(), "P")
        src = _dev(sources[f], "source", plan.image_dtype)
        keep += [P, src]
        Pp[f], Sp[f] = P.data_ptr(), src.data_ptr()
        if want_sample:
            samples.append(torch.empty(B, H, W, 2, dtype=torch.float32, device=dev))
            samp[f] = samples[-1].data_ptr()
        if want_color:
            colors.append(torch.empty(B, 3, H, W, dtype=torch.float32, device=dev))
            colp[f] = colors[-1].data_ptr()
    depth = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev) if want_depth else None
    check(plan.lib.vsl_warp_forward(ctypes.byref(plan.desc), scale_index, disp.data_ptr(), iK.data_ptr(),
                                    ctypes.byref(Pp), ctypes.byref(Sp), ptr(depth), ctypes.byref(samp),
                                    ctypes.byref(colp), _stream()), "vsl_warp_forward")
    return depth, samples, colors


# --------------------------------------------------------------------------------------------------
# stand-alone layers
# --------------------------------------------------------------------------------------------------
class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K, arith):
        depth, inv_K = _dev(depth, "depth"), _dev(inv_K, "inv_K")
        B, _, H, W = depth.shape
        cam = torch.empty(B, 4, H * W, dtype=torch.float32, device=depth.device)
        check(_lib.load().vsl_backproject_forward(B, H, W, arith, depth.data_ptr(), inv_K.data_ptr(),
                                                  cam.data_ptr(), _stream()), "vsl_backproject_forward")
        ctx.save_for_backward(inv_K)
        ctx.shape = (B, H, W)
        return cam

    @staticmethod
    def backward(ctx, g):
        (inv_K,) = ctx.saved_tensors
        B, H, W = ctx.shape
        g = _dev(g, "grad")
        gd = torch.empty(B, 1, H, W, dtype=torch.float32, device=g.device)
        check(_lib.load().vsl_backproject_backward(B, H, W, g.data_ptr(), inv_K.data_ptr(), gd.data_ptr(),
                                                   _stream()), "vsl_backproject_backward")
        return gd, None, None


def backproject(depth, inv_K, arith="auto"):
    if arith == "auto":
        arith = calibrate_arith(depth.shape[0], depth.shape[2], depth.shape[3], depth.device)
    return _Backproject.apply(depth, inv_K, arith)


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, P, height, width, eps, arith):
        points, P = _dev(points, "points"), _dev(P, "P")
        B = points.shape[0]
        pix = torch.empty(B, height, width, 2, dtype=torch.float32, device=points.device)
        check(_lib.load().vsl_project_forward(B, height, width, eps, arith, points.data_ptr(), P.data_ptr(),
                                              pix.data_ptr(), _stream()), "vsl_project_forward")
        ctx.save_for_backward(points, P)
        ctx.meta = (B, height, width, eps)
        return pix

    @staticmethod
    def backward(ctx, g):
        points, P = ctx.saved_tensors
        B, H, W, eps = ctx.meta
        lib = _lib.load()
        g = _dev(g, "grad")
        gpts = torch.empty_like(points)
        gP = torch.empty_like(P)
        nbytes = lib.vsl_project_workspace_bytes(B, H, W)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=g.device)
        check(lib.vsl_project_backward(B, H, W, eps, points.data_ptr(), P.data_ptr(), g.data_ptr(),
                                       gpts.data_ptr(), gP.data_ptr(), ws.data_ptr(), nbytes, _stream()),
              "vsl_project_backward")
        return gpts, gP, None, None, None, None


def project(points, P, height, width, eps=1e-7, arith="auto"):
    if arith == "auto":
        arith = calibrate_arith(points.shape[0], height, width, points.device)
    return _Project.apply(points, P, height, width, float(np.float32(eps)), arith)


class _Ssim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _dev(x, "x"), _dev(y, "y")
        B, C, H, W = x.shape
        out = torch.empty_like(x)
        check(_lib.load().vsl_ssim_forward(B, C, H, W, x.data_ptr(), y.data_ptr(), out.data_ptr(), _stream()),
              "vsl_ssim_forward")
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        B, C, H, W = x.shape
        g = _dev(g, "grad")
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        lib = _lib.load()
        nbytes = lib.vsl_ssim_workspace_bytes(B, C, H, W)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=g.device)
        check(lib.vsl_ssim_backward(B, C, H, W, x.data_ptr(), y.data_ptr(), g.data_ptr(), ptr(gx), ptr(gy),
                                    ws.data_ptr(), nbytes, _stream()), "vsl_ssim_backward")
        return gx, gy


def ssim(x, y):
    return _Ssim.apply(x, y)


class _ReprojLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, no_ssim, arith):
        pred, target = _dev(pred, "pred"), _dev(target, "target")
        B, C, H, W = pred.shape
        if C != 3:
            raise ValueError("compute_reprojection_loss expects 3-channel images")
        out = torch.empty(B, 1, H, W, dtype=torch.float32, device=pred.device)
        check(_lib.load().vsl_reprojection_loss_forward(B, H, W, int(no_ssim), arith, pred.data_ptr(),
                                                        target.data_ptr(), out.data_ptr(), _stream()),
              "vsl_reprojection_loss_forward")
        ctx.save_for_backward(pred, target)
        ctx.no_ssim = int(no_ssim)
        return out

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        B, _, H, W = pred.shape
        g = _dev(g, "grad")
        gp = torch.empty_like(pred) if ctx.needs_input_grad[0] else None
        gt = torch.empty_like(target) if ctx.needs_input_grad[1] else None
        lib = _lib.load()
        nbytes = 0 if ctx.no_ssim else lib.vsl_ssim_workspace_bytes(B, 3, H, W)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=g.device) if nbytes else None
        check(lib.vsl_reprojection_loss_backward(B, H, W, ctx.no_ssim, pred.data_ptr(), target.data_ptr(),
                                                 g.data_ptr(), ptr(gp), ptr(gt), ptr(ws), nbytes, _stream()),
              "vsl_reprojection_loss_backward")
        return gp, gt, None, None


def reprojection_loss(pred, target, no_ssim=False, arith=0):
    return _ReprojLoss.apply(pred, target, no_ssim, arith)


class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img):
        disp, img = _dev(disp, "disp"), _dev(img, "img")
        B, _, H, W = disp.shape
        lib = _lib.load()
        nbytes = lib.vsl_smooth_workspace_bytes(B, H, W)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=disp.device)
        loss = torch.empty((), dtype=torch.float32, device=disp.device)
        check(lib.vsl_smooth_loss_forward(B, H, W, disp.data_ptr(), img.data_ptr(), loss.data_ptr(), ws.data_ptr(),
                                          nbytes, _stream()), "vsl_smooth_loss_forward")
        ctx.save_for_backward(disp, img)
        return loss

    @staticmethod
    def backward(ctx, g):
        disp, img = ctx.saved_tensors
        B, _, H, W = disp.shape
        g = _dev(g, "grad")
        gd = torch.empty_like(disp)
        check(_lib.load().vsl_smooth_loss_backward(B, H, W, disp.data_ptr(), img.data_ptr(), g.data_ptr(),
                                                   gd.data_ptr(), _stream()), "vsl_smooth_loss_backward")
        return gd, None


def smooth_loss(disp, img):
    return _SmoothLoss.apply(disp, img)


# --------------------------------------------------------------------------------------------------
# depth metrics and the GAN prior's loss
# --------------------------------------------------------------------------------------------------
def _metrics_ws(dev):
    lib = _lib.load()
    nbytes = lib.vsl_metrics_workspace_bytes()
    return torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev), nbytes


def depth_errors(gt, pred):
    """compute_depth_errors (layers.py:335-353) in one kernel: returns the seven metrics as a [7] device tensor
    (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3)."""
    gt, pred = _dev(gt.detach().reshape(-1), "gt"), _dev(pred.detach().reshape(-1), "pred")
    if gt.numel() != pred.numel() or gt.numel() == 0:
        raise ValueError("gt and pred must have the same, non-zero number of elements")
    out = torch.empty(7, dtype=torch.float32, device=gt.device)
    ws, nbytes = _metrics_ws(gt.device)
    check(_lib.load().vsl_depth_errors(gt.numel(), gt.data_ptr(), pred.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes,
                                       _stream()), "vsl_depth_errors")
    return out


def depth_losses(depth_pred, depth_gt, crop=(153, 371, 44, 1197), clamp=(1e-3, 80.0)):
    """Trainer.compute_depth_losses (trainer.py:688-716) on the device: [7] tensor of the metrics."""
    depth_pred, depth_gt = _dev(depth_pred.detach(), "depth_pred"), _dev(depth_gt.detach(), "depth_gt")
    B, _, h, w = depth_pred.shape
    Bg, _, gh, gw = depth_gt.shape
    if B != Bg:
        raise ValueError("depth_pred and depth_gt have different batch sizes")
    out = torch.empty(7, dtype=torch.float32, device=depth_pred.device)
    ws, nbytes = _metrics_ws(depth_pred.device)
    c = (ctypes.c_int * 4)(*[int(v) for v in crop])
    check(_lib.load().vsl_depth_losses(B, h, w, gh, gw, ctypes.byref(c), float(np.float32(clamp[0])), float(np.float32(clamp[1])),
                                       depth_pred.data_ptr(), depth_gt.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes,
                                       _stream()), "vsl_depth_losses")
    return out


class _SLlog(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fake, real):
        fake, real = _dev(fake, "fake"), _dev(real, "real")
        if fake.shape != real.shape:
            raise ValueError("SLlog needs equal shapes (the reference's resize branch reads an unassigned name)")
        dev = fake.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        stats = torch.empty(3, dtype=torch.float32, device=dev)
        ws, nbytes = _metrics_ws(dev)
        check(_lib.load().vsl_sllog_forward(fake.numel(), fake.data_ptr(), real.data_ptr(), loss.data_ptr(), stats.data_ptr(),
                                            ws.data_ptr(), nbytes, _stream()), "vsl_sllog_forward")
        ctx.save_for_backward(fake, real, stats)
        return loss

    @staticmethod
    def backward(ctx, g):
        fake, real, stats = ctx.saved_tensors
        g = _dev(g, "grad")
        gf = torch.empty_like(fake) if ctx.needs_input_grad[0] else None
        gr = torch.empty_like(real) if ctx.needs_input_grad[1] else None
        check(_lib.load().vsl_sllog_backward(fake.numel(), fake.data_ptr(), real.data_ptr(), stats.data_ptr(), g.data_ptr(),
                                             ptr(gf), ptr(gr), _stream()), "vsl_sllog_backward")
        return gf, gr


def sllog(fake, real):
    return _SLlog.apply(fake, real)
