"""CPU tests of the image-pyramid oracle (oracle/pil_pyramid_oracle.py) and of the library's host-side
coefficient tables: pinned to Pillow through the committed goldens (tests/golden/pyramid/make_golden_pyramid.py)."""
import os
import sys

import numpy as np
import pytest

from oracle import pil_pyramid_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "pyramid", "pyramid_pil.npz")
NAMES = ["iid_64x96", "smooth_96x160", "edges_32x64"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_pillow_goldens(name):
    g = np.load(GOLDEN)
    levels, tensors = O.pyramid(g[name + "/u8_0"], 4)
    for s in range(1, 4):
        assert np.array_equal(levels[s], g["%s/u8_%d" % (name, s)]), (name, s)
    assert np.array_equal(tensors[3], g[name + "/f32_3"])
    assert tensors[0].dtype == np.float32 and tensors[0].shape == (3,) + g[name + "/u8_0"].shape[:2]


def test_oracle_against_live_pillow_if_present():
    """Not needed for the pin (the goldens are), but cheap where Pillow is installed."""
    pil = pytest.importorskip("PIL.Image")
    rng = np.random.RandomState(7)
    img = rng.randint(0, 256, (40, 72, 3)).astype(np.uint8)
    ref = np.asarray(pil.fromarray(img, "RGB").resize((36, 20), pil.LANCZOS))
    assert np.array_equal(O.resize_lanczos(img, 20, 36), ref)


def test_batched_oracle_equals_per_image():
    rng = np.random.RandomState(3)
    batch = rng.randint(0, 256, (3, 16, 32, 3)).astype(np.uint8)
    lv, ts = O.pyramid(batch, 3)
    for b in range(3):
        lv1, ts1 = O.pyramid(batch[b], 3)
        for s in range(3):
            assert np.array_equal(lv[s][b], lv1[s])
            assert np.array_equal(ts[s][b], ts1[s])


@pytest.mark.parametrize("n", [640, 192, 320, 256, 1024, 96, 48, 24, 12, 4, 2])
def test_library_coefficient_tables_equal_pillow_restatement(n):
    """vsl_pyramid_coefficients is host-only C++ (no device call): same tables as the oracle, hence as Pillow."""
    from unsupervised_pose_estimation_b200.input_pipeline import pyramid_coefficients
    bounds, coefs = pyramid_coefficients(n, n // 2)
    k, ob, oc = O.precompute_coeffs(n, n // 2)
    assert k <= 13
    assert np.array_equal(bounds, ob)
    assert np.array_equal(coefs[:, :k], oc) and not coefs[:, k:].any()
    assert (coefs.sum(1) - (1 << O.PRECISION_BITS)).__abs__().max() <= 13  # rows sum to one up to rounding


def test_pyramid_needs_cuda():
    import torch
    from unsupervised_pose_estimation_b200 import _lib
    from unsupervised_pose_estimation_b200.input_pipeline import FramePyramid
    with pytest.raises(_lib.VslError):
        FramePyramid(1, 32, 64, 4, device="cpu")
    if not torch.cuda.is_available():
        desc = _lib.VslPyramidDesc(_lib.VSL_ABI_VERSION, 1, 30, 64, 4, 0)   # 30 is not a multiple of 8
        assert _lib.load().vsl_pyramid_workspace_bytes(desc) == 0


def test_oracle_reproduces_the_level0_resize_goldens():
    """resize_lanczos at arbitrary ratios (the decoded file image -> level 0, datasets/mono_dataset2.py:85-89, :107-109)
    against outputs of the reference's own transforms.Resize(..., LANCZOS) on PIL images."""
    here = os.path.join(os.path.dirname(__file__), "golden", "pyramid")
    if here not in sys.path:
        sys.path.insert(0, here)
    import resize_cases as gen
    g = np.load(os.path.join(here, "resize_pil.npz"))
    for name, h, w, oh, ow, family in gen.CASES:
        img = gen.make_input(name, h, w, family)
        assert np.array_equal(O.resize_lanczos(img, oh, ow), g[name]), name
