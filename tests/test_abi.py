"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/vsl.h declares, and
validates its arguments before touching CUDA (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from unsupervised_pose_estimation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vsl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vsl_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for sym in header_symbols():
        assert hasattr(lib, sym), sym
    assert lib.vsl_abi_version() == _lib.VSL_ABI_VERSION
    assert lib.vsl_status_string(0) == b"ok"
    assert lib.vsl_status_string(-4) == b"unsupported option"


def make_desc(**kw):
    d = _lib.VslDesc()
    d.abi_version = _lib.VSL_ABI_VERSION
    d.batch, d.height, d.width = 2, 64, 96
    d.num_scales = 4
    for i in range(4):
        d.scale_ids[i] = i
    d.num_src = 2
    d.flags = _lib.FLAG_AUTOMASK
    d.min_disp, d.disp_range, d.eps, d.smooth_weight = 0.01, 9.99, 1e-7, 1e-3
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def test_descriptor_validation_without_gpu():
    lib = _lib.load()
    assert lib.vsl_loss_workspace_bytes(ctypes.byref(make_desc())) > 0
    assert lib.vsl_loss_workspace_bytes(ctypes.byref(make_desc(abi_version=99))) == 0
    assert lib.vsl_loss_workspace_bytes(ctypes.byref(make_desc(height=60))) == 0  # 60 >> 3 not exact
    assert lib.vsl_loss_workspace_bytes(ctypes.byref(make_desc(num_src=9))) == 0
    buf = _lib.VslLossBuffers()
    ws = ctypes.create_string_buffer(64)
    rc = lib.vsl_loss_forward_backward(ctypes.byref(make_desc(abi_version=0)), ctypes.byref(buf), ws, 64, None)
    assert rc == -1
    rc = lib.vsl_loss_forward_backward(ctypes.byref(make_desc()), None, ws, 64, None)
    assert rc == -2
    rc = lib.vsl_loss_forward_backward(ctypes.byref(make_desc(flags=_lib.FLAG_AUTOMASK | _lib.FLAG_V1_MULTISCALE)),
                                       ctypes.byref(buf), ws, 64, None)
    assert rc == -4
    rc = lib.vsl_loss_forward_backward(ctypes.byref(make_desc()), ctypes.byref(buf), ws, 64, None)
    assert rc == -5  # workspace too small is reported before any pointer is dereferenced
    assert lib.vsl_ssim_forward(1, 3, 1, 8, None, None, None, None) == -1
    assert lib.vsl_ssim_forward(1, 3, 8, 8, None, None, None, None) == -2
    assert lib.vsl_project_workspace_bytes(2, 64, 96) > 0
    assert lib.vsl_smooth_workspace_bytes(2, 64, 96) > 0


def test_product_refuses_cpu():
    import torch
    from unsupervised_pose_estimation_b200 import functional as VF
    from unsupervised_pose_estimation_b200.trainer import LossPath, make_opt
    with pytest.raises(_lib.VslError):
        LossPath(make_opt(), device="cpu")
    with pytest.raises(_lib.VslError):
        VF.ssim(torch.rand(1, 3, 8, 8), torch.rand(1, 3, 8, 8))
