"""GPU parity of the on-device colour augmentation (vsl_color_aug_forward) against the Pillow/torchvision-pinned
oracle and the goldens made by the reference's own transform objects: byte-exact 8-bit images, bit-exact tensors."""
import os

import numpy as np
import pytest
import torch

from oracle import color_aug_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "augment", "color_aug.npz")


def _run(batch_u8, params, dtype=torch.float32):
    from unsupervised_pose_estimation_b200.input_pipeline import ColorAugment
    B, H, W, _ = batch_u8.shape
    aug = ColorAugment(B, H, W, "cuda", dtype)
    out, u8 = aug(torch.from_numpy(batch_u8).cuda(), params, want_u8=True)
    torch.cuda.synchronize()
    return out.cpu(), u8.cpu().numpy()


def _golden():
    z = np.load(GOLDEN)
    n = len([k for k in z.files if k.endswith("/in")])
    imgs, params, want = [], [], []
    for i in range(n):
        key = "case%03d" % i
        f = z[key + "/factors"]
        imgs.append(z[key + "/in"])
        want.append(z[key + "/out_u8"])
        params.append(dict(order=[int(v) for v in z[key + "/order"]], brightness=float(f[0]), contrast=float(f[1]),
                           saturation=float(f[2]), hue=float(f[3]), flip=bool(z[key + "/flags"][0]),
                           autocontrast=bool(z[key + "/flags"][1])))
    return z, np.stack(imgs), params, np.stack(want)


def test_color_aug_equals_the_reference_transforms():
    z, imgs, params, want = _golden()
    out, u8 = _run(imgs, params)
    bad = [i for i in range(len(params)) if not np.array_equal(u8[i], want[i])]
    assert not bad, (bad[:5], params[bad[0]])
    for i in range(8):
        assert np.array_equal(out[i].numpy(), z["case%03d/out_f32" % i]), i
    for i in range(len(params)):
        assert np.array_equal(out[i].numpy(), O.to_tensor(want[i])), i


@pytest.mark.parametrize("shape", [(12, 192, 640), (12, 96, 320), (3, 24, 80), (2, 7, 5)])
def test_color_aug_equals_oracle_at_the_pyramid_shapes(shape):
    from unsupervised_pose_estimation_b200.input_pipeline import draw_color_aug_params
    B, H, W = shape
    rng = np.random.RandomState(H + W)
    yy, xx = np.mgrid[0:H, 0:W]
    batch = np.empty((B, H, W, 3), np.uint8)
    for b in range(B):
        if b % 3 == 0:
            batch[b] = rng.randint(0, 256, (H, W, 3))
        elif b % 3 == 1:  # smooth, narrow range: autocontrast stretches it
            batch[b] = np.stack([110 + 30 * np.sin(xx / 17.0 + c + b) * np.cos(yy / 11.0) for c in range(3)], -1)
        else:             # saturating
            batch[b] = np.clip(rng.normal(200, 90, (H, W, 3)), 0, 255)
    torch.manual_seed(17 + H)
    params = [None if b == 4 else draw_color_aug_params() for b in range(B)]
    # wider factors than the reference's (0.8, 1.2): both blend branches and the clipping
    if B > 1:
        params[1] = dict(order=[2, 0, 3, 1], brightness=1.7, contrast=0.3, saturation=1.9, hue=-0.5, flip=True, autocontrast=True)
    out, u8 = _run(batch, params)
    for b in range(B):
        ref = batch[b] if params[b] is None else O.color_aug(batch[b], params[b])
        assert np.array_equal(u8[b], ref), (b, params[b])
        assert np.array_equal(out[b].numpy(), O.to_tensor(ref)), b


def test_hue_round_trip_is_exact_for_every_rgb_triple():
    v = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], -1).astype(np.uint8).reshape(1, 4096, 4096, 3)
    for hue in (0.0, 0.1, -0.37):
        prm = dict(order=[3, 0, 1, 2], brightness=1.0, contrast=1.0, saturation=1.0, hue=hue, flip=False, autocontrast=False)
        out, u8 = _run(rgb, [prm])
        assert np.array_equal(u8[0], O.adjust_hue(rgb[0], hue)), hue


def test_color_aug_bf16_output_and_errors():
    from unsupervised_pose_estimation_b200 import _lib
    from unsupervised_pose_estimation_b200.input_pipeline import ColorAugment, draw_color_aug_params
    rng = np.random.RandomState(5)
    batch = rng.randint(0, 256, (2, 16, 24, 3)).astype(np.uint8)
    torch.manual_seed(3)
    params = [draw_color_aug_params(), draw_color_aug_params()]
    f32, _ = _run(batch, params)
    b16, _ = _run(batch, params, torch.bfloat16)
    assert torch.equal(b16, f32.bfloat16())
    with pytest.raises(_lib.VslError):
        ColorAugment(2, 16, 24, "cpu")
    aug = ColorAugment(2, 16, 24)
    with pytest.raises(_lib.VslError):
        aug(torch.from_numpy(batch), params)            # host tensor
    with pytest.raises(ValueError):
        aug(torch.from_numpy(batch).cuda(), params[:1])  # one record per image
    with pytest.raises(ValueError):
        aug(torch.from_numpy(batch).cuda(), [dict(params[0], hue=0.7), params[1]])
