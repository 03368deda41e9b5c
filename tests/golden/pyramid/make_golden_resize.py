"""Golden vectors for the arbitrary-ratio resize of the decoded file image to level 0, made by the reference's own
transform object (build container: Pillow + torchvision):

    python tests/golden/pyramid/make_golden_resize.py   ->  tests/golden/pyramid/resize_pil.npz

`self.resize[0] = transforms.Resize((height, width), interpolation=Image.ANTIALIAS)` applied to
`inputs[(n, im, -1)]`, the image as loaded from disk (datasets/mono_dataset2.py:85-89, :107-109).  Inputs are
regenerated from their seed (`make_input`), only the outputs are stored.
"""
import os
import sys

import numpy as np
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from resize_cases import CASES, make_input  # noqa: E402


def main():
    out = {}
    for name, h, w, oh, ow, family in CASES:
        img = make_input(name, h, w, family)
        res = transforms.Resize((oh, ow), interpolation=Image.LANCZOS)(Image.fromarray(img, "RGB"))
        out[name] = np.asarray(res)
        assert out[name].shape == (oh, ow, 3)
    import PIL
    out["pillow_version"] = np.array(PIL.__version__)
    np.savez_compressed(os.path.join(HERE, "resize_pil.npz"), **out)
    print("wrote resize_pil.npz", {k: v.shape for k, v in out.items() if k != "pillow_version"})


if __name__ == "__main__":
    main()
