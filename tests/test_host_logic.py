"""Host-side logic that needs no GPU: synthetic factory, pose helpers, option gating, bench model."""
import importlib.util
import os

import pytest
import torch

from oracle import vsl_oracle as O
from unsupervised_pose_estimation_b200 import layers as L
from unsupervised_pose_estimation_b200 import synthetic
from unsupervised_pose_estimation_b200.trainer import ViewSynthesisLossMixin, make_opt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_is_deterministic_and_well_formed():
    a = synthetic.make_batch(2, 32, 64, [0, -1, 1, "s"], seed=5)
    b = synthetic.make_batch(2, 32, 64, [0, -1, 1, "s"], seed=5)
    for k in a[0]:
        assert torch.equal(a[0][k], b[0][k])
    for k in a[2]:
        assert torch.equal(a[2][k], b[2][k])
    inputs = a[0]
    assert inputs[("color", "s", 2)].shape == (2, 3, 8, 16)
    assert inputs[("K", 1)][0, 0, 0].item() == pytest.approx(0.58 * 32)
    assert torch.allclose(inputs[("K", 0)][0] @ inputs[("inv_K", 0)][0], torch.eye(4), atol=1e-5)
    assert inputs["stereo_T"][0, 0, 3].item() == pytest.approx(0.1)


@pytest.mark.parametrize("invert", [False, True])
def test_pose_helpers_match_oracle(invert):
    g = torch.Generator().manual_seed(0)
    aa, tr = 0.01 * torch.randn(5, 1, 3, generator=g), 0.01 * torch.randn(5, 1, 3, generator=g)
    assert torch.equal(L.transformation_from_parameters(aa, tr, invert), O.transformation_from_parameters(aa, tr, invert))
    d = torch.rand(2, 1, 4, 4, generator=g)
    for a, b in zip(L.disp_to_depth(d, 0.1, 100), O.disp_to_depth(d, 0.1, 100)):
        assert torch.equal(a, b)


def test_layers_reexports_reference_names():
    for name in ["SLlog", "RMSE_log", "depth_to_disp", "disp_to_depth", "transformation_from_parameters",
                 "get_translation_matrix", "rot_from_axisangle", "ConvBlock", "batchNorm", "Conv3x3",
                 "BackprojectDepth", "Project3D", "upsample", "deconv", "get_smooth_loss", "SSIM",
                 "compute_depth_errors"]:
        assert hasattr(L, name), name
    assert L.BackprojectDepth(2, 4, 6).batch_size == 2 and L.Project3D(2, 4, 6).eps == 1e-7
    x = torch.rand(1, 3, 8, 8)
    assert L.ConvBlock(3, 4)(x).shape == (1, 4, 8, 8) and L.upsample(x).shape == (1, 3, 16, 16)


def test_unsupported_flags_fail_loudly():
    """posecnn with a stereo frame is the one combination the path refuses (the reference itself fails there)."""
    class P(ViewSynthesisLossMixin):
        pass
    p = P()
    p.opt = make_opt(pose_model_type="posecnn", frame_ids=[0, -1, 1, "s"])
    with pytest.raises(NotImplementedError):
        p._vsl_plan()


def test_bench_byte_model():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.algorithmic_bytes_per_pixel(2) == pytest.approx(50.5625)
    assert bench.algorithmic_bytes_per_pixel(3, 2) == pytest.approx(36.59375)  # C3, bf16 images


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the reference's CPU path, oracle port) runs without a GPU and prints ONE JSON
    line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "px/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("C2")


def test_input_pipeline_classes_have_no_cpu_path():
    """FrameResize / FramePyramid / ColorAugment are CUDA only: asking for a CPU device fails loudly (before any
    library call), like the loss path itself."""
    from unsupervised_pose_estimation_b200 import _lib
    from unsupervised_pose_estimation_b200.input_pipeline import ColorAugment, FramePyramid, FrameResize
    with pytest.raises(_lib.VslError):
        FrameResize(1, 20, 30, 16, 24, device="cpu")
    with pytest.raises(_lib.VslError):
        FramePyramid(1, 16, 24, 2, device="cpu")
    with pytest.raises(_lib.VslError):
        ColorAugment(1, 16, 24, device="cpu")


def test_resize_workspace_and_argument_checks():
    """vsl_resize_* argument validation through the C ABI (host only: no kernel is launched)."""
    import ctypes
    from unsupervised_pose_estimation_b200 import _lib
    lib = _lib.load()
    assert lib.vsl_resize_workspace_bytes(0, 10, 10, 5, 5) == 0
    assert lib.vsl_resize_workspace_bytes(1, 10, 10, 0, 5) == 0
    n = lib.vsl_resize_workspace_bytes(2, 375, 1242, 192, 640)
    # tables (two axes) + the horizontal pass's output [2, 375, 640, 3]
    assert n >= 2 * 375 * 640 * 3 and n % 256 == 0
    assert lib.vsl_resize_forward(2, 375, 1242, 192, 640, None, None, None, n, None) == -2      # null pointers
    assert lib.vsl_color_aug_workspace_bytes(0, 4, 4) == 0
    assert lib.vsl_color_aug_workspace_bytes(3, 8, 8) >= 3 * 8 * 8 * 3
    assert lib.vsl_color_aug_forward(1, 8, 8, 7, ctypes.c_void_p(256), ctypes.c_void_p(256), ctypes.c_void_p(256), None,
                                     ctypes.c_void_p(256), 1 << 20, None) == -4                 # unsupported dtype
