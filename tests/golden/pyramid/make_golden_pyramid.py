"""Golden vectors for the image-pyramid preprocessing, made by the reference's own transform objects.

Run in the build container (needs Pillow + torchvision, like the reference's dataset):
    python tests/golden/pyramid/make_golden_pyramid.py

Follows datasets/mono_dataset2.py:85-89 (``transforms.Resize((H // s, W // s), interpolation=Image.ANTIALIAS)``
per level, each applied to the previous level, :103-117) and ``transforms.ToTensor()``.  ``Image.ANTIALIAS`` was
removed in Pillow 10; it was an alias of ``Image.LANCZOS``, which is used here.
"""
import os

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_pyramid(img_u8, num_scales=4):
    """[H, W, 3] uint8 -> (levels uint8 HWC, tensors float32 CHW) via PIL / torchvision."""
    h, w = img_u8.shape[:2]
    interp = Image.LANCZOS
    resize = {i: transforms.Resize((h // 2 ** i, w // 2 ** i), interpolation=interp) for i in range(num_scales)}
    to_tensor = transforms.ToTensor()
    level = {-1: Image.fromarray(img_u8, "RGB")}
    for i in range(num_scales):
        level[i] = resize[i](level[i - 1])   # level 0: same size -> PIL returns a copy
    return ([np.asarray(level[i]) for i in range(num_scales)],
            [to_tensor(level[i]).numpy() for i in range(num_scales)])


def make_images():
    rng = np.random.RandomState(1234)
    imgs = {}
    imgs["iid_64x96"] = rng.randint(0, 256, (64, 96, 3)).astype(np.uint8)
    yy, xx = np.mgrid[0:96, 0:160]
    smooth = np.stack([127.5 + 127.5 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0 - c) for c in range(3)], -1)
    imgs["smooth_96x160"] = np.clip(smooth + rng.randn(96, 160, 3) * 4, 0, 255).astype(np.uint8)
    edge = np.zeros((32, 64, 3), np.uint8)   # saturating step edges: exercises the 8-bit clip of both passes
    edge[:, 20:41] = 255
    edge[10:20] = 255 - edge[10:20]
    imgs["edges_32x64"] = edge
    return imgs


if __name__ == "__main__":
    out = {}
    for name, img in make_images().items():
        levels, tensors = reference_pyramid(img)
        out[name + "/u8_0"] = img
        for s in range(1, 4):
            out["%s/u8_%d" % (name, s)] = levels[s]
        out[name + "/f32_3"] = tensors[3]
        assert np.array_equal(levels[0], img)
        assert np.array_equal(tensors[0], np.moveaxis(img, -1, 0).astype(np.float32) / np.float32(255))
    import PIL
    out["pillow_version"] = np.array(PIL.__version__)
    np.savez_compressed(os.path.join(HERE, "pyramid_pil.npz"), **out)
    print("wrote pyramid_pil.npz (Pillow %s)" % PIL.__version__)
