"""Golden vectors for the colour augmentation of the ``color_aug`` inputs, made by the reference's own transform
objects (build container: Pillow + torchvision):

    python tests/golden/augment/make_golden_color_aug.py   ->  tests/golden/augment/color_aug.npz

The reference builds ``transforms.Compose([ColorJitter((0.8,1.2),(0.8,1.2),(0.8,1.2),(-0.1,0.1)),
RandomHorizontalFlip(p=0.5), RandomAutocontrast()])`` (datasets/mono_dataset2.py:71-97) and calls it on every 8-bit
PIL level (``self.to_tensor(color_aug(f))``, :124).  Each call draws fresh parameters from torch's global generator.
Here the SAME objects are called on PIL images under a fixed seed; the draws are recorded by re-running the generator
from the same state through ``oracle.color_aug_oracle.draw_params`` (which restates the order of the draws), and the
outputs (8-bit image and ToTensor result) are committed.  Cases cover every permutation of the four jitter
operations several times, both flip / autocontrast outcomes, and constant / two-level / saturating images.
"""
import os
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)
from oracle import color_aug_oracle as O  # noqa: E402


def make_images():
    rng = np.random.default_rng(11)
    h, w = 32, 48
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = np.stack([127 + 100 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0 - c) for c in range(3)], -1)
    imgs = {
        "iid": rng.integers(0, 256, (h, w, 3)),
        "smooth": smooth,
        "dark": rng.integers(0, 40, (h, w, 3)),
        "bright": rng.integers(200, 256, (h, w, 3)),
        "narrow": rng.integers(100, 131, (h, w, 3)),
        "constant": np.full((h, w, 3), 77),
        "gray": np.repeat(rng.integers(0, 256, (h, w, 1)), 3, -1),
        "primaries": np.stack([(xx * 7) % 256, (yy * 11) % 256, ((xx + yy) * 5) % 256], -1),
    }
    return {k: np.clip(v, 0, 255).astype(np.uint8) for k, v in imgs.items()}


def main():
    aug = transforms.Compose([
        transforms.ColorJitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1)),
        transforms.RandomHorizontalFlip(p=0.5),
        transforms.RandomAutocontrast()])
    to_tensor = transforms.ToTensor()
    out = {}
    imgs = make_images()
    n = 0
    for rep in range(12):
        for name, img in imgs.items():
            torch.manual_seed(1000 + n)
            state = torch.get_rng_state()
            res = aug(Image.fromarray(img, "RGB"))
            torch.set_rng_state(state)
            prm = O.draw_params()
            key = "case%03d" % n
            out[key + "/in"] = img
            out[key + "/out_u8"] = np.asarray(res)
            if n < 8:  # ToTensor is byte / 255 for every case; eight cases pin it
                out[key + "/out_f32"] = to_tensor(res).numpy()
            out[key + "/order"] = np.array(prm["order"], np.int32)
            out[key + "/factors"] = np.array([prm["brightness"], prm["contrast"], prm["saturation"], prm["hue"]], np.float64)
            out[key + "/flags"] = np.array([prm["flip"], prm["autocontrast"]], np.int32)
            n += 1
    import PIL
    import torchvision
    out["versions"] = np.array("pillow %s torchvision %s" % (PIL.__version__, torchvision.__version__))
    np.savez_compressed(os.path.join(HERE, "color_aug.npz"), **out)
    print("wrote color_aug.npz with %d cases" % n)


if __name__ == "__main__":
    main()
