"""The cases of tests/golden/pyramid/resize_pil.npz: shapes and seeded inputs (numpy only, so the GPU tests can
regenerate the inputs without Pillow / torchvision).  Outputs: make_golden_resize.py."""
import numpy as np

# (name, native h, native w, level-0 h, level-0 w, family)
CASES = [
    ("kitti_375x1242_to_192x640", 375, 1242, 192, 640, "smooth"),      # BASELINE configs 1, 3, 5
    ("scared_1024x1280_to_256x320", 1024, 1280, 256, 320, "smooth"),   # config 2: ratio 4 (25 taps per axis)
    ("ratio_1p2_120x154_to_100x128", 120, 154, 100, 128, "iid"),       # config 4's ratio (375x1242 -> 320x1024)
    ("odd_101x149_to_64x96", 101, 149, 64, 96, "iid"),
    ("upscale_40x60_to_64x96", 40, 60, 64, 96, "iid"),                 # scale < 1: support stays 3
    ("same_width_200x96_to_64x96", 200, 96, 64, 96, "edges"),          # horizontal pass skipped
    ("same_height_64x300_to_64x96", 64, 300, 64, 96, "edges"),         # vertical pass skipped
]


def make_input(name, h, w, family):
    rng = np.random.RandomState(h * 10007 + w)
    if family == "iid":
        return rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    if family == "smooth":
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([127.5 + 127.5 * np.sin(xx / 23.0 + c) * np.cos(yy / 17.0 - c) for c in range(3)], -1)
        return np.clip(img + rng.randn(h, w, 3) * 6, 0, 255).astype(np.uint8)
    img = np.zeros((h, w, 3), np.uint8)       # saturating step edges
    img[:, w // 3: 2 * w // 3] = 255
    img[h // 4: h // 2] = 255 - img[h // 4: h // 2]
    return img
