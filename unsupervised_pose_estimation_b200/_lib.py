"""ctypes binding of libvsl_b200.so (C ABI declared in include/vsl.h).

The library is the only compute path: if it cannot be loaded the package raises, it never
falls back to PyTorch ops or to the CPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p

VSL_ABI_VERSION = 3
VSL_MAX_SCALES = 4
VSL_MAX_SRC = 4

FLAG_AUTOMASK = 1 << 0
FLAG_AVG_REPROJECTION = 1 << 1
FLAG_NO_SSIM = 1 << 2
FLAG_V1_MULTISCALE = 1 << 3
FLAG_FORWARD_ONLY = 1 << 4

DTYPE_F32 = 0
DTYPE_BF16 = 1

ARITH_TRUE_DIV = 1 << 0
ARITH_DOT_NOFMA = 1 << 1
ARITH_DOT_REVERSE = 1 << 2
ARITH_UPS_RIGHT = 1 << 3
ARITH_UPS_NOFMA = 1 << 4
ARITH_TAP_NOFMA = 1 << 5
ARITH_MEAN_DIV = 1 << 6
ARITH_DOT3_NOFMA = 1 << 7
ARITH_DOT3_REVERSE = 1 << 8
ARITH_DOTKT_NOFMA = 1 << 9
ARITH_DOTKT_REVERSE = 1 << 10
ARITH_NORM_SEQ = 1 << 11

# every symbol include/vsl.h declares (tests check the shared object exports all of them)
EXPORTED_SYMBOLS = [
    "vsl_abi_version", "vsl_status_string", "vsl_last_cuda_error",
    "vsl_loss_workspace_bytes", "vsl_loss_workspace_init", "vsl_loss_forward_backward", "vsl_loss_forward_backward_timed",
    "vsl_event_create", "vsl_event_destroy", "vsl_event_elapsed_ms", "vsl_loss_combine_grads",
    "vsl_warp_forward", "vsl_probe_bmm", "vsl_pose_forward", "vsl_pose_backward",
    "vsl_backproject_forward", "vsl_backproject_backward",
    "vsl_project_forward", "vsl_project_workspace_bytes", "vsl_project_backward",
    "vsl_ssim_forward", "vsl_ssim_workspace_bytes", "vsl_ssim_backward",
    "vsl_reprojection_loss_forward", "vsl_reprojection_loss_backward",
    "vsl_smooth_workspace_bytes", "vsl_smooth_loss_forward", "vsl_smooth_loss_backward",
    "vsl_pyramid_workspace_bytes", "vsl_pyramid_plan", "vsl_pyramid_forward", "vsl_pyramid_coefficients",
    "vsl_source_grad_upstream", "vsl_grid_sample_backward_source",
    "vsl_pyramid_forward_flip", "vsl_stereo_transform",
    "vsl_posecnn_workspace_bytes", "vsl_posecnn_forward", "vsl_posecnn_backward",
    "vsl_metrics_workspace_bytes", "vsl_depth_errors", "vsl_depth_losses", "vsl_sllog_forward", "vsl_sllog_backward",
    "vsl_color_aug_workspace_bytes", "vsl_color_aug_forward",
    "vsl_resize_workspace_bytes", "vsl_resize_plan", "vsl_resize_forward",
]


class VslDesc(Structure):
    _fields_ = [
        ("abi_version", c_int32), ("batch", c_int32), ("height", c_int32), ("width", c_int32),
        ("num_scales", c_int32), ("scale_ids", c_int32 * VSL_MAX_SCALES), ("num_src", c_int32),
        ("flags", c_int32), ("image_dtype", c_int32), ("arith", c_int32),
        ("min_disp", c_float), ("disp_range", c_float), ("eps", c_float), ("smooth_weight", c_float),
        ("smooth_level_bias", c_int32),
    ]


class VslLossBuffers(Structure):
    _fields_ = [
        ("target", c_void_p * VSL_MAX_SCALES), ("source", c_void_p * VSL_MAX_SRC),
        ("disp", c_void_p * VSL_MAX_SCALES), ("inv_K", c_void_p), ("P", c_void_p * VSL_MAX_SRC),
        ("K", c_void_p), ("T", c_void_p * VSL_MAX_SRC), ("T_scale", (c_void_p * VSL_MAX_SRC) * VSL_MAX_SCALES),
        ("noise", c_void_p * VSL_MAX_SCALES), ("predictive_mask", c_void_p * VSL_MAX_SCALES),
        ("losses", c_void_p), ("mask", c_void_p * VSL_MAX_SCALES),
        ("grad_disp_photo", c_void_p * VSL_MAX_SCALES), ("grad_disp_smooth", c_void_p * VSL_MAX_SCALES),
        ("smooth_norm", c_void_p), ("grad_P", c_void_p), ("grad_predictive_mask", c_void_p * VSL_MAX_SCALES),
        ("side_depth", c_void_p * VSL_MAX_SCALES), ("side_sample", (c_void_p * VSL_MAX_SRC) * VSL_MAX_SCALES),
        ("side_color", (c_void_p * VSL_MAX_SRC) * VSL_MAX_SCALES),
        ("winner", c_void_p * VSL_MAX_SCALES),
    ]


class VslPyramidDesc(Structure):
    _fields_ = [
        ("abi_version", c_int32), ("batch", c_int32), ("height", c_int32), ("width", c_int32),
        ("num_levels", c_int32), ("out_dtype", c_int32),
    ]


class VslAugParams(Structure):
    _fields_ = [
        ("order", c_int32 * 4), ("factor", c_float * 3), ("hue_shift", c_int32), ("flip", c_int32),
        ("autocontrast", c_int32), ("enabled", c_int32), ("reserved", c_int32),
    ]


class VslError(RuntimeError):
    pass


_LIB = None


def lib_path():
    """The in-tree library; VSL_LIB_PATH points the loader at another build of the same ABI (developer knob
    for measuring kernel variants side by side — never a different compute path)."""
    return os.environ.get("VSL_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvsl_b200.so")


def load():
    """Load (building first if the in-tree .so is missing or stale and nvcc is present)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    from . import build as _build
    # one rank builds, the others wait on the lock and then find a complete, renamed-into-place file
    with _build.build_lock():
        if not os.environ.get("VSL_LIB_PATH") and _build.is_stale() and _build.find_nvcc() is not None:
            _build.build()
        if not os.path.exists(path):
            raise VslError(
                "libvsl_b200.so is missing and nvcc is not available to build it; run "
                "`python -m unsupervised_pose_estimation_b200.build`. There is no CPU/PyTorch fallback.")
        lib = ctypes.CDLL(path)
    if lib.vsl_abi_version() != VSL_ABI_VERSION:
        raise VslError("libvsl_b200.so ABI version mismatch; rebuild it")

    vp = c_void_p
    lib.vsl_status_string.restype = c_char_p
    lib.vsl_status_string.argtypes = [c_int]
    lib.vsl_loss_workspace_bytes.restype = c_size_t
    lib.vsl_loss_workspace_bytes.argtypes = [POINTER(VslDesc)]
    lib.vsl_loss_workspace_init.argtypes = [POINTER(VslDesc), vp, c_size_t, vp]
    lib.vsl_loss_forward_backward.argtypes = [POINTER(VslDesc), POINTER(VslLossBuffers), vp, c_size_t, vp]
    lib.vsl_loss_forward_backward_timed.argtypes = [POINTER(VslDesc), POINTER(VslLossBuffers), vp, c_size_t, vp, vp, vp]
    lib.vsl_event_create.argtypes = [POINTER(c_void_p)]
    lib.vsl_event_destroy.argtypes = [vp]
    lib.vsl_event_elapsed_ms.argtypes = [vp, vp, POINTER(c_float)]
    lib.vsl_loss_combine_grads.argtypes = [POINTER(VslDesc), vp, POINTER(VslLossBuffers),
                                           POINTER(c_void_p * VSL_MAX_SCALES), vp, vp, vp]
    lib.vsl_warp_forward.argtypes = [POINTER(VslDesc), c_int, vp, vp, POINTER(c_void_p * VSL_MAX_SRC),
                                     POINTER(c_void_p * VSL_MAX_SRC), vp, POINTER(c_void_p * VSL_MAX_SRC),
                                     POINTER(c_void_p * VSL_MAX_SRC), vp]
    lib.vsl_pose_forward.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp]
    lib.vsl_pose_backward.argtypes = [c_int, c_int, vp, vp, vp, vp, vp, vp]
    lib.vsl_probe_bmm.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp]
    lib.vsl_backproject_forward.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp]
    lib.vsl_backproject_backward.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp]
    lib.vsl_project_forward.argtypes = [c_int, c_int, c_int, c_float, c_int, vp, vp, vp, vp]
    lib.vsl_project_workspace_bytes.restype = c_size_t
    lib.vsl_project_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.vsl_project_backward.argtypes = [c_int, c_int, c_int, c_float, vp, vp, vp, vp, vp, vp, c_size_t, vp]
    lib.vsl_ssim_forward.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp]
    lib.vsl_ssim_workspace_bytes.restype = c_size_t
    lib.vsl_ssim_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.vsl_ssim_backward.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, c_size_t, vp]
    lib.vsl_reprojection_loss_forward.argtypes = [c_int, c_int, c_int, c_int, c_int, vp, vp, vp, vp]
    lib.vsl_reprojection_loss_backward.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, c_size_t, vp]
    lib.vsl_smooth_workspace_bytes.restype = c_size_t
    lib.vsl_smooth_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.vsl_smooth_loss_forward.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, c_size_t, vp]
    lib.vsl_smooth_loss_backward.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, vp]
    lib.vsl_pyramid_workspace_bytes.restype = c_size_t
    lib.vsl_pyramid_workspace_bytes.argtypes = [POINTER(VslPyramidDesc)]
    lib.vsl_pyramid_plan.argtypes = [POINTER(VslPyramidDesc), vp, c_size_t, vp]
    lib.vsl_pyramid_forward.argtypes = [POINTER(VslPyramidDesc), vp, POINTER(c_void_p * VSL_MAX_SCALES),
                                        POINTER(c_void_p * VSL_MAX_SCALES), vp, c_size_t, vp]
    lib.vsl_pyramid_coefficients.argtypes = [c_int, c_int, POINTER(c_int32), POINTER(c_int32), c_int]
    lib.vsl_pyramid_forward_flip.argtypes = [POINTER(VslPyramidDesc), vp, vp, POINTER(c_void_p * VSL_MAX_SCALES),
                                             POINTER(c_void_p * VSL_MAX_SCALES), vp, c_size_t, vp]
    lib.vsl_stereo_transform.argtypes = [c_int, vp, vp, c_float, vp, vp]
    lib.vsl_source_grad_upstream.argtypes = [POINTER(VslDesc), vp, POINTER(c_void_p * VSL_MAX_SCALES), vp,
                                             POINTER(c_void_p * VSL_MAX_SCALES), vp]
    lib.vsl_grid_sample_backward_source.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp]
    lib.vsl_posecnn_workspace_bytes.restype = c_size_t
    lib.vsl_posecnn_workspace_bytes.argtypes = [POINTER(VslDesc)]
    lib.vsl_posecnn_forward.argtypes = [POINTER(VslDesc), POINTER(c_void_p * VSL_MAX_SCALES), c_int,
                                        POINTER(c_void_p * VSL_MAX_SRC), POINTER(c_void_p * VSL_MAX_SRC),
                                        POINTER(c_int32 * VSL_MAX_SRC), c_int, vp, vp, vp, c_size_t, vp]
    lib.vsl_posecnn_backward.argtypes = [POINTER(VslDesc), c_int, POINTER(c_void_p * VSL_MAX_SRC),
                                         POINTER(c_void_p * VSL_MAX_SRC), POINTER(c_int32 * VSL_MAX_SRC), vp, vp,
                                         POINTER(c_void_p * VSL_MAX_SRC), POINTER(c_void_p * VSL_MAX_SRC), vp, vp]
    lib.vsl_metrics_workspace_bytes.restype = c_size_t
    lib.vsl_metrics_workspace_bytes.argtypes = []
    lib.vsl_depth_errors.argtypes = [c_size_t, vp, vp, vp, vp, c_size_t, vp]
    lib.vsl_depth_losses.argtypes = [c_int, c_int, c_int, c_int, c_int, POINTER(c_int * 4), c_float, c_float, vp, vp, vp,
                                     vp, c_size_t, vp]
    lib.vsl_sllog_forward.argtypes = [c_size_t, vp, vp, vp, vp, vp, c_size_t, vp]
    lib.vsl_sllog_backward.argtypes = [c_size_t, vp, vp, vp, vp, vp, vp, vp]
    lib.vsl_color_aug_workspace_bytes.restype = c_size_t
    lib.vsl_color_aug_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.vsl_color_aug_forward.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, c_size_t, vp]
    lib.vsl_resize_workspace_bytes.restype = c_size_t
    lib.vsl_resize_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.vsl_resize_plan.argtypes = [c_int, c_int, c_int, c_int, c_int, vp, c_size_t, vp]
    lib.vsl_resize_forward.argtypes = [c_int, c_int, c_int, c_int, c_int, vp, vp, vp, c_size_t, vp]
    _LIB = lib
    return lib


def check(status, what):
    if status != 0:
        lib = load()
        msg = lib.vsl_status_string(status).decode()
        extra = ""
        if status == -6:
            extra = " (cudaError %d)" % lib.vsl_last_cuda_error()
        raise VslError("%s failed: %s%s" % (what, msg, extra))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
