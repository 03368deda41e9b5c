"""The colour-augmentation oracle (oracle/color_aug_oracle.py) against the goldens made by the reference's own
transform objects, and -- where Pillow is installed -- exhaustively against the live library."""
import os

import numpy as np
import pytest

from oracle import color_aug_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "augment", "color_aug.npz")


def golden_cases():
    z = np.load(GOLDEN)
    n = len([k for k in z.files if k.endswith("/in")])
    for i in range(n):
        key = "case%03d" % i
        f = z[key + "/factors"]
        prm = dict(order=[int(v) for v in z[key + "/order"]], brightness=float(f[0]), contrast=float(f[1]),
                   saturation=float(f[2]), hue=float(f[3]), flip=bool(z[key + "/flags"][0]),
                   autocontrast=bool(z[key + "/flags"][1]))
        yield key, z, prm


def test_oracle_reproduces_the_reference_transforms():
    orders = set()
    n = 0
    for key, z, prm in golden_cases():
        got = O.color_aug(z[key + "/in"], prm)
        assert np.array_equal(got, z[key + "/out_u8"]), (key, prm)
        if key + "/out_f32" in z.files:
            assert np.array_equal(O.to_tensor(got), z[key + "/out_f32"]), key
        orders.add(tuple(prm["order"]))
        n += 1
    assert n == 96 and len(orders) >= 20   # nearly every permutation of the four operations occurs


def test_draw_params_consumes_the_generator_like_the_reference():
    torch = pytest.importorskip("torch")
    tv = pytest.importorskip("torchvision")
    from torchvision import transforms
    torch.manual_seed(5)
    a = O.draw_params()
    b = O.draw_params()
    torch.manual_seed(5)
    fn_idx, br, c, s, h = transforms.ColorJitter.get_params((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))
    flip = bool(torch.rand(1) < 0.5)
    auto = bool(torch.rand(1).item() < 0.5)
    assert a == dict(order=[int(v) for v in fn_idx], brightness=br, contrast=c, saturation=s, hue=h, flip=flip,
                     autocontrast=auto)
    fn_idx2 = transforms.ColorJitter.get_params((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))[0]
    assert b["order"] == [int(v) for v in fn_idx2]


def test_pixel_maps_equal_live_pillow_exhaustively():
    Image = pytest.importorskip("PIL.Image")
    v = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    im = Image.fromarray(rgb, "RGB")
    assert np.array_equal(np.asarray(im.convert("HSV")), O.rgb_to_hsv(rgb))
    assert np.array_equal(np.asarray(Image.fromarray(rgb, "HSV").convert("RGB")), O.hsv_to_rgb(rgb))
    assert np.array_equal(np.asarray(im.convert("L")), O.to_gray(rgb))
    a = np.arange(256, dtype=np.uint8)
    i1, i2 = np.repeat(a[:, None], 256, 1), np.repeat(a[None, :], 256, 0)
    rng = np.random.default_rng(0)
    for alpha in list(rng.uniform(0.5, 1.5, 24)) + [0.0, 1.0, 0.8, 1.2, 1.7, -0.2]:
        ref = np.asarray(Image.blend(Image.fromarray(i1, "L"), Image.fromarray(i2, "L"), alpha))
        assert np.array_equal(ref, O.blend(i1, i2, alpha)), alpha


def test_autocontrast_equals_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    from PIL import ImageOps
    rng = np.random.default_rng(3)
    for lo, hi in ((0, 255), (3, 250), (100, 130), (77, 77), (0, 1), (254, 255), (17, 201)):
        img = rng.integers(lo, hi + 1, (24, 40, 3)).astype(np.uint8)
        img[0, 0], img[0, 1] = lo, hi
        assert np.array_equal(np.asarray(ImageOps.autocontrast(Image.fromarray(img, "RGB"))), O.autocontrast(img))


def test_host_side_draws_equal_the_oracle_restatement():
    """input_pipeline.draw_color_aug_params (product, host-side bookkeeping) and the oracle's draw_params consume the
    generator identically."""
    torch = pytest.importorskip("torch")
    from unsupervised_pose_estimation_b200.input_pipeline import draw_color_aug_params
    torch.manual_seed(9)
    a = [draw_color_aug_params() for _ in range(5)]
    torch.manual_seed(9)
    b = [O.draw_params() for _ in range(5)]
    assert a == b
