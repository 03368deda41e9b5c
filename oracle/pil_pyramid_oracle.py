"""TEST-ONLY oracle of the reference's image-pyramid preprocessing.

The reference builds ``("color", f, s)`` on the CPU in its dataset (datasets/mono_dataset2.py:103-124):
level s is ``transforms.Resize((H // 2**s, W // 2**s), interpolation=Image.ANTIALIAS)`` of level s-1 (a PIL
image, 8 bits per channel), and every level then goes through ``transforms.ToTensor()`` (HWC uint8 ->
CHW float32 / 255).  ``Image.ANTIALIAS`` is Pillow's LANCZOS filter.

The arithmetic lives in a third-party dependency that is not vendored under /root/reference: **Pillow**
(the reference pins no version; this image has Pillow 12.2.0).  This file restates Pillow's published
8-bit resampling algorithm (src/libImaging/Resample.c: ``precompute_coeffs``, ``normalize_coeffs_8bpc``,
``ImagingResampleHorizontal_8bpc`` / ``Vertical_8bpc``, two passes, horizontal first, each pass rounded to
8 bits) in numpy.  It is pinned against Pillow itself: tests/golden/pyramid/make_golden_pyramid.py runs the
reference's own transform objects here and commits the results; tests/test_pyramid_oracle.py requires this
restatement to reproduce them byte for byte.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos(x):
    """Pillow's lanczos_filter: truncated sinc, support 3 (Resample.c)."""
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size).

    Returns (ksize, bounds[out_size, 2] = (first input index, tap count), coefs[out_size, ksize] int32)."""
    scale = float(in_size) / float(out_size)
    filterscale = max(scale, 1.0)
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coefs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)   # C cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = [lanczos((x + xmin - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            coefs[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, n)
    return ksize, bounds, coefs


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis(img, out_size, axis):
    """One 8-bit pass along ``axis`` of an [..., H, W, C] uint8 array (axis = -3 rows / -2 columns)."""
    in_size = img.shape[axis]
    _, bounds, coefs = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        lo, n = bounds[xx]
        k = coefs[xx, :n].astype(np.int64)
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(k, src[lo:lo + n], axes=(0, 0))
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_lanczos(img, out_h, out_w):
    """PIL ``Image.resize((out_w, out_h), Image.LANCZOS)`` of an [..., H, W, C] uint8 array."""
    h, w = img.shape[-3], img.shape[-2]
    if out_w != w:
        img = resample_axis(img, out_w, -2)   # horizontal pass first
    if out_h != h:
        img = resample_axis(img, out_h, -3)
    return img


def to_tensor(img):
    """transforms.ToTensor(): [..., H, W, C] uint8 -> [..., C, H, W] float32, value / 255 (IEEE division)."""
    return (np.moveaxis(img, -1, -3).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def pyramid(frame_u8, num_levels):
    """MonoDataset.preprocess for one frame: level 0 is the frame itself, level s = resize(level s-1).

    Returns (levels_u8 [list of [..., h, w, 3] uint8], tensors [list of [..., 3, h, w] float32])."""
    h, w = frame_u8.shape[-3], frame_u8.shape[-2]
    levels = [frame_u8]
    for s in range(1, num_levels):
        levels.append(resize_lanczos(levels[-1], h // (2 ** s), w // (2 ** s)))
    return levels, [to_tensor(l) for l in levels]
