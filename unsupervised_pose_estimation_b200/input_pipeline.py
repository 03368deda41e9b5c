"""On-GPU input pipeline of the loss path: 8-bit frames in, the ``("color", f, s)`` pyramid out.

The reference prepares its colour inputs on the CPU, inside ``MonoDataset`` (datasets/mono_dataset2.py):
``preprocess`` (:103-124) resizes every frame to the four pyramid levels with
``transforms.Resize(..., interpolation=Image.ANTIALIAS)`` (each level from the previous one, :85-89), runs
``transforms.ToTensor()`` on each, and ``Trainer.process_batch`` then copies the fp32 tensors to the device
(trainer.py:373-374).  For the loss path that is 59 MB per step at config 1 and the step becomes PCIe-bound.

``FramePyramid`` moves that work behind the copy: the host hands over the level-0 frames as uint8 HWC
(``np.asarray(pil_image)`` — 13 MB per step at config 1) and the GPU reproduces Pillow's 8-bit LANCZOS
resampling and the ``/255`` conversion bit for bit (``vsl_pyramid_forward``, csrc/vsl_input.cu), writing the
same ``("color", f, s)`` tensors the reference's dataset would have produced.  CUDA only, like the rest of the
package: there is no CPU path.

    pyr = FramePyramid(batch=12, height=192, width=640, num_levels=4, device="cuda")
    levels = pyr(frame_u8)                      # list of [B,3,H>>s,W>>s] float32 tensors (static buffers)
    inputs = preprocess({0: u8_0, -1: u8_m1, 1: u8_p1}, pyramids)   # the reference's key layout
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import VSL_ABI_VERSION, VSL_MAX_SCALES, VslAugParams, VslPyramidDesc, check


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class FramePyramid:
    """One frame slot: ``num_levels`` static output tensors + the workspace (coefficient tables, 8-bit levels).

    ``levels`` selects which tensors are written (default: all); the loss path needs every level of the target
    frame but only level 0 of the source frames (trainer.py:502, :534-537).
    """

    def __init__(self, batch, height, width, num_levels=4, device="cuda", dtype=torch.float32, levels=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.VslError("FramePyramid runs on CUDA only (the reference's CPU pyramid is its own dataset code)")
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be float32 or bfloat16")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.batch, self.height, self.width, self.num_levels = batch, height, width, num_levels
        self.wanted = sorted(range(num_levels) if levels is None else levels)
        last = max(self.wanted) + 1   # levels past the last wanted one are never formed
        self.desc = VslPyramidDesc(VSL_ABI_VERSION, batch, height, width, last,
                                   _lib.DTYPE_BF16 if dtype == torch.bfloat16 else _lib.DTYPE_F32)
        self.ws_bytes = int(self.lib.vsl_pyramid_workspace_bytes(ctypes.byref(self.desc)))
        if self.ws_bytes == 0:
            raise ValueError("bad pyramid shape: %dx%d must be multiples of 2^%d" % (height, width, last - 1))
        with torch.cuda.device(self.device):
            self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=self.device)
            off = (-self.ws.data_ptr()) % 256
            self.ws_ptr = self.ws.data_ptr() + off
            self.out = {s: torch.empty(batch, 3, height >> s, width >> s, dtype=dtype, device=self.device)
                        for s in self.wanted}
            check(self.lib.vsl_pyramid_plan(ctypes.byref(self.desc), ctypes.c_void_p(self.ws_ptr), self.ws_bytes,
                                            _stream()), "vsl_pyramid_plan")
            torch.cuda.current_stream().synchronize()  # the tables come from temporary host memory

    def __call__(self, frame_u8, want_u8=False, flip=None):
        """frame_u8: [B,H,W,3] uint8 CUDA tensor.  Returns {level: tensor}; with want_u8 also the 8-bit levels.
        ``flip``: optional [B] uint8 CUDA tensor, non-zero = mirror that image left-right first (the dataset's
        ``do_flip``, datasets/mono_dataset2.py:151-156)."""
        if flip is not None and (flip.device != self.device or flip.dtype != torch.uint8 or flip.numel() != self.batch
                                 or not flip.is_contiguous()):
            raise _lib.VslError("flip must be a contiguous uint8 tensor of %d flags on %s" % (self.batch, self.device))
        if frame_u8.device != self.device or frame_u8.dtype != torch.uint8 or not frame_u8.is_contiguous():
            raise _lib.VslError("frames must be contiguous uint8 tensors on %s (got %s %s)"
                                % (self.device, frame_u8.dtype, frame_u8.device))
        if tuple(frame_u8.shape) != (self.batch, self.height, self.width, 3):
            raise ValueError("frame has shape %s, expected %s"
                             % (tuple(frame_u8.shape), (self.batch, self.height, self.width, 3)))
        lv = (ctypes.c_void_p * VSL_MAX_SCALES)()
        for s, t in self.out.items():
            lv[s] = t.data_ptr()
        u8 = None
        u8p = None
        if want_u8:
            u8 = {s: torch.empty(self.batch, self.height >> s, self.width >> s, 3, dtype=torch.uint8, device=self.device)
                  for s in range(1, self.desc.num_levels)}
            u8p = (ctypes.c_void_p * VSL_MAX_SCALES)()
            for s, t in u8.items():
                u8p[s] = t.data_ptr()
        check(self.lib.vsl_pyramid_forward_flip(ctypes.byref(self.desc), ctypes.c_void_p(frame_u8.data_ptr()),
                                                ctypes.c_void_p(flip.data_ptr()) if flip is not None else None,
                                                ctypes.byref(lv), ctypes.byref(u8p) if u8p is not None else None,
                                                ctypes.c_void_p(self.ws_ptr), self.ws_bytes, _stream()),
              "vsl_pyramid_forward_flip")
        return (self.out, u8) if want_u8 else self.out


class FrameResize:
    """``self.resize[0]`` of ``MonoDataset`` (datasets/mono_dataset2.py:85-89, :107-109) for a batch on the GPU: the
    decoded 8-bit file images at their native resolution -> level 0 at (height, width), byte for byte what
    ``transforms.Resize((height, width), interpolation=Image.ANTIALIAS)`` gives on the PIL images.  The result is the
    uint8 HWC batch ``FramePyramid`` takes.  Optional: shipping native-resolution frames costs more host->device
    bytes than shipping level 0 (3.8x at KITTI's 1242 x 375), so it only pays where the host cannot resize.

        to_level0 = FrameResize(batch=12, in_height=375, in_width=1242, height=192, width=640)
        frame_u8 = to_level0(native_u8)          # [B,192,640,3] uint8 (static buffer)
    """

    def __init__(self, batch, in_height, in_width, height, width, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.VslError("FrameResize runs on CUDA only (the reference's CPU resize is its own dataset code)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.shape = (batch, in_height, in_width, height, width)
        self.ws_bytes = int(self.lib.vsl_resize_workspace_bytes(*self.shape))
        if self.ws_bytes == 0:
            raise ValueError("bad resize shape %s" % (self.shape,))
        with torch.cuda.device(self.device):
            self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=self.device)
            self.ws_ptr = self.ws.data_ptr() + (-self.ws.data_ptr()) % 256
            self.out = torch.empty(batch, height, width, 3, dtype=torch.uint8, device=self.device)
            check(self.lib.vsl_resize_plan(*self.shape, ctypes.c_void_p(self.ws_ptr), self.ws_bytes, _stream()),
                  "vsl_resize_plan")
            torch.cuda.current_stream().synchronize()  # the tables come from temporary host memory

    def __call__(self, native_u8):
        batch, in_h, in_w, _, _ = self.shape
        if native_u8.device != self.device or native_u8.dtype != torch.uint8 or not native_u8.is_contiguous():
            raise _lib.VslError("frames must be contiguous uint8 tensors on %s (got %s %s)"
                                % (self.device, native_u8.dtype, native_u8.device))
        if tuple(native_u8.shape) != (batch, in_h, in_w, 3):
            raise ValueError("frame has shape %s, expected %s" % (tuple(native_u8.shape), (batch, in_h, in_w, 3)))
        check(self.lib.vsl_resize_forward(*self.shape, ctypes.c_void_p(native_u8.data_ptr()),
                                          ctypes.c_void_p(self.out.data_ptr()), ctypes.c_void_p(self.ws_ptr),
                                          self.ws_bytes, _stream()), "vsl_resize_forward")
        return self.out


class LossInputPipeline:
    """``MonoDataset.preprocess`` + the host->device copy for the frames the loss path reads.

    ``frame_ids`` as in the options (first entry is the target, trainer.py:502); the target frame gets every
    level, the source frames level 0 only (``all_levels=True`` forms every level of every frame, like the
    reference's dataset does for the networks' benefit).
    """

    def __init__(self, opt, device="cuda", dtype=torch.float32, all_levels=False):
        self.frame_ids = list(opt.frame_ids)
        self.opt = opt
        self.device = torch.device(device)
        self._intrinsics = {}
        n = len(opt.scales)
        self.pyramids = {
            f: FramePyramid(opt.batch_size, opt.height, opt.width, n, device, dtype,
                            levels=None if (all_levels or i == 0) else [0])
            for i, f in enumerate(self.frame_ids)}

    def __call__(self, frames_u8, inputs=None, flip=None, side_left=None):
        """frames_u8: {frame_id: [B,H,W,3] uint8 CUDA tensor}.  Fills / returns ``inputs[("color", f, s)]``.

        ``flip`` ([B] uint8 on the device): the items' ``do_flip`` draws — every frame of a flagged item is mirrored
        before the pyramid is formed, and with a stereo frame ``inputs["stereo_T"]`` gets the matching baseline sign
        (datasets/mono_dataset2.py:155-156, :197-203; ``side_left``: [B] uint8, 1 where the item is a left image)."""
        inputs = {} if inputs is None else inputs
        for f in self.frame_ids:
            for s, t in self.pyramids[f](frames_u8[f], flip=flip).items():
                inputs[("color", f, s)] = t
        if "s" in self.frame_ids and (flip is not None or side_left is not None):
            inputs["stereo_T"] = self.stereo_T(flip, side_left)
        return inputs

    def stereo_T(self, flip=None, side_left=None, baseline=0.1):
        """inputs["stereo_T"] [B,4,4] (datasets/mono_dataset2.py:197-203) from the per-item flags, on the device."""
        B = self.opt.batch_size
        dev = next(iter(self.pyramids.values())).device
        T = torch.empty(B, 4, 4, dtype=torch.float32, device=dev)
        lib = _lib.load()
        check(lib.vsl_stereo_transform(B, ctypes.c_void_p(flip.data_ptr()) if flip is not None else None,
                               ctypes.c_void_p(side_left.data_ptr()) if side_left is not None else None,
                               float(baseline), ctypes.c_void_p(T.data_ptr()), _stream()), "vsl_stereo_transform")
        return T

    def intrinsics(self, K_norm, inputs=None):
        """inputs[("K", s)] / [("inv_K", s)] [B,4,4] for every scale (datasets/mono_dataset2.py:168-177), as PLAN
        CONSTANTS: the reference re-derives them for every item in its DataLoader workers and copies 8 matrices per
        image to the device every step, although ``self.K`` is one constant per dataset.  Here they are computed once
        per normalised K with the reference's own numpy operations (row scaling by ``width // 2**s`` /
        ``height // 2**s``, ``np.linalg.pinv``: the bit pattern of inv_K feeds the bit-exact ray computation, so it has
        to be numpy's) and stay resident; later calls return the cached device tensors."""
        import numpy as np
        K_norm = np.asarray(K_norm, dtype=np.float32)
        key = K_norm.tobytes()
        if key not in self._intrinsics:
            opt, dev = self.opt, next(iter(self.pyramids.values())).device
            out = {}
            for s in range(len(opt.scales)):
                K = K_norm.copy()
                K[0, :] *= opt.width // (2 ** s)
                K[1, :] *= opt.height // (2 ** s)
                inv_K = np.linalg.pinv(K)
                out[("K", s)] = torch.from_numpy(K).unsqueeze(0).repeat(opt.batch_size, 1, 1).to(dev)
                out[("inv_K", s)] = torch.from_numpy(inv_K).unsqueeze(0).repeat(opt.batch_size, 1, 1).to(dev)
            self._intrinsics[key] = out
        if inputs is not None:
            inputs.update(self._intrinsics[key])
        return self._intrinsics[key]


def draw_color_aug_params(brightness=(0.8, 1.2), contrast=(0.8, 1.2), saturation=(0.8, 1.2), hue=(-0.1, 0.1),
                          p_flip=0.5, p_autocontrast=0.5):
    """The random draws of ONE call of the reference's ``transforms_aug`` (datasets/mono_dataset2.py:92-97), taken
    from torch's global generator in the order torchvision takes them: ``ColorJitter.get_params`` (``randperm(4)``,
    then one ``uniform_`` each for brightness, contrast, saturation, hue), ``RandomHorizontalFlip`` (``rand(1) < p``),
    ``RandomAutocontrast`` (``rand(1) < p``).  Host-side bookkeeping; the pixels are processed by ColorAugment."""
    order = [int(v) for v in torch.randperm(4)]
    b = float(torch.empty(1).uniform_(brightness[0], brightness[1]))
    c = float(torch.empty(1).uniform_(contrast[0], contrast[1]))
    s = float(torch.empty(1).uniform_(saturation[0], saturation[1]))
    h = float(torch.empty(1).uniform_(hue[0], hue[1]))
    flip = bool(torch.rand(1) < p_flip)
    auto = bool(torch.rand(1).item() < p_autocontrast)
    return dict(order=order, brightness=b, contrast=c, saturation=s, hue=h, flip=flip, autocontrast=auto)


class ColorAugment:
    """``self.to_tensor(color_aug(f))`` of ``MonoDataset.preprocess`` (datasets/mono_dataset2.py:124) for a batch of
    equally sized 8-bit images on the GPU: ColorJitter + RandomHorizontalFlip + RandomAutocontrast + ToTensor, byte
    for byte what torchvision / Pillow produce on the PIL images, given each image's draws.

        aug = ColorAugment(batch=12, height=192, width=640)
        params = [draw_color_aug_params() if do_color_aug else None for _ in range(12)]   # None: ToTensor only
        color_aug = aug(frame_u8, params)            # [B,3,H,W] float32 (static buffer)

    The reference calls the transform once per (frame, level), so every level has its own draws; use one instance per
    level size.  CUDA only."""

    def __init__(self, batch, height, width, device="cuda", dtype=torch.float32):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.VslError("ColorAugment runs on CUDA only (the reference's CPU augmentation is its own dataset code)")
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("dtype must be float32 or bfloat16")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.batch, self.height, self.width, self.dtype = batch, height, width, dtype
        self.ws_bytes = int(self.lib.vsl_color_aug_workspace_bytes(batch, height, width))
        if self.ws_bytes == 0:
            raise ValueError("bad shape %dx%dx%d" % (batch, height, width))
        with torch.cuda.device(self.device):
            self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=self.device)
            self.ws_ptr = self.ws.data_ptr() + (-self.ws.data_ptr()) % 256
            self.out = torch.empty(batch, 3, height, width, dtype=dtype, device=self.device)
            self.params_dev = torch.empty(batch * ctypes.sizeof(VslAugParams), dtype=torch.uint8, device=self.device)
        self.params_host = torch.empty(batch * ctypes.sizeof(VslAugParams), dtype=torch.uint8).pin_memory()

    def pack_params(self, params):
        """list of per-image dicts (draw_color_aug_params) or None -> the pinned host array of VslAugParams"""
        import numpy as np
        if len(params) != self.batch:
            raise ValueError("need %d parameter records, got %d" % (self.batch, len(params)))
        arr = (VslAugParams * self.batch).from_buffer(self.params_host.numpy())
        for rec, prm in zip(arr, params):
            if prm is None:
                rec.enabled = 0
                rec.order[:] = [0, 1, 2, 3]
                rec.factor[:] = [1.0, 1.0, 1.0]
                rec.hue_shift = rec.flip = rec.autocontrast = rec.reserved = 0
                continue
            if sorted(prm["order"]) != [0, 1, 2, 3]:
                raise ValueError("order must be a permutation of 0..3")
            if not -0.5 <= prm["hue"] <= 0.5:
                raise ValueError("hue_factor (%r) is not in [-0.5, 0.5]." % (prm["hue"],))  # torchvision's check
            rec.enabled = 1
            rec.order[:] = [int(v) for v in prm["order"]]
            rec.factor[:] = [float(prm["brightness"]), float(prm["contrast"]), float(prm["saturation"])]
            rec.hue_shift = int(np.int32(prm["hue"] * 255).astype(np.uint8))   # _functional_pil.adjust_hue
            rec.flip = int(bool(prm["flip"]))
            rec.autocontrast = int(bool(prm["autocontrast"]))
            rec.reserved = 0
        return self.params_host

    def __call__(self, frame_u8, params, want_u8=False):
        """frame_u8: [B,H,W,3] uint8 CUDA tensor; params: see pack_params.  Returns the float tensor (static
        buffer), with want_u8 also the augmented 8-bit images."""
        if frame_u8.device != self.device or frame_u8.dtype != torch.uint8 or not frame_u8.is_contiguous():
            raise _lib.VslError("frames must be contiguous uint8 tensors on %s (got %s %s)"
                                % (self.device, frame_u8.dtype, frame_u8.device))
        if tuple(frame_u8.shape) != (self.batch, self.height, self.width, 3):
            raise ValueError("frame has shape %s, expected %s"
                             % (tuple(frame_u8.shape), (self.batch, self.height, self.width, 3)))
        self.params_dev.copy_(self.pack_params(params), non_blocking=True)
        u8 = torch.empty_like(frame_u8) if want_u8 else None
        check(self.lib.vsl_color_aug_forward(self.batch, self.height, self.width,
                                             _lib.DTYPE_BF16 if self.dtype == torch.bfloat16 else _lib.DTYPE_F32,
                                             ctypes.c_void_p(frame_u8.data_ptr()), ctypes.c_void_p(self.params_dev.data_ptr()),
                                             ctypes.c_void_p(self.out.data_ptr()),
                                             ctypes.c_void_p(u8.data_ptr()) if u8 is not None else None,
                                             ctypes.c_void_p(self.ws_ptr), self.ws_bytes, _stream()),
              "vsl_color_aug_forward")
        return (self.out, u8) if want_u8 else self.out


def pyramid_coefficients(in_size, out_size):
    """Host-only: the integer coefficient table of one axis as the library computes it (bounds, coefs)."""
    import numpy as np
    lib = _lib.load()
    bounds = np.zeros((out_size, 2), np.int32)
    coefs = np.zeros((out_size, 13), np.int32)
    check(lib.vsl_pyramid_coefficients(in_size, out_size, bounds.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                       coefs.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 13),
          "vsl_pyramid_coefficients")
    return bounds, coefs
