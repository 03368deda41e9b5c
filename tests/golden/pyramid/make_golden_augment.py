"""Golden vectors for the dataset-side AUGMENTATIONS that now run on the GPU, made by the reference's own objects
(build container: Pillow + torchvision):

    python tests/golden/pyramid/make_golden_augment.py   ->  tests/golden/pyramid/augment_pil.npz

flip:  MonoDataset.get_color's `color.transpose(Image.FLIP_LEFT_RIGHT)` (datasets/kitti_dataset.py via
       mono_dataset2.py:151-156) followed by the pyramid of mono_dataset2.py:85-89, :103-117 and ToTensor.
"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_pyramid import make_images, reference_pyramid  # noqa: E402


def main():
    out = {}
    for name, img in make_images().items():
        flipped = np.asarray(Image.fromarray(img, "RGB").transpose(Image.FLIP_LEFT_RIGHT))
        levels, tensors = reference_pyramid(flipped)
        for s in range(4):
            out["flip/%s/u8_%d" % (name, s)] = levels[s]
        out["flip/%s/f32_3" % name] = tensors[3]
    import PIL
    out["pillow_version"] = np.array(PIL.__version__)
    np.savez_compressed(os.path.join(HERE, "augment_pil.npz"), **out)
    print("wrote augment_pil.npz", sorted(out)[:6])


if __name__ == "__main__":
    main()
