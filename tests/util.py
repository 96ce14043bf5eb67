"""Shared helpers for the tests: fixtures on disk and seeded synthetic inputs."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
W512_PATH = os.path.join(GOLDEN, "w_512.dat")
PNG512_PATH = os.path.join(GOLDEN, "512.png")


def load_512_rgb():
    """(3, 512, 512) float32 0..255 planar — af::loadImage(path, true) (main.cpp:153)."""
    from PIL import Image
    a = np.asarray(Image.open(PNG512_PATH).convert("RGB"), np.float32)
    return np.ascontiguousarray(a.transpose(2, 0, 1))


def load_512_gray(oracle):
    """gray f32 (non-integer values) via the reference's weights (main.cpp:154)."""
    return oracle.rgb2gray(load_512_rgb())


def load_w512():
    return np.fromfile(W512_PATH, np.float32).reshape(512, 512)


def natural_image(rows, cols, seed=0, integer=False):
    """Natural-statistics test image of any size: the 512.png luma, mirrored-tiled, shifted and lightly
    noised (white noise would make Rx well-conditioned and hide the hard cases)."""
    from PIL import Image
    g = np.asarray(Image.open(PNG512_PATH).convert("RGB"), np.float32)
    g = 0.299 * g[..., 0] + 0.587 * g[..., 1] + 0.114 * g[..., 2]
    g = np.concatenate([g, g[:, ::-1]], 1)
    g = np.concatenate([g, g[::-1, :]], 0)  # 1024 x 1024, seamless
    rng = np.random.default_rng(seed)
    oy, ox = rng.integers(0, 1024, 2)
    yy = (np.arange(rows) + oy) % 1024
    xx = (np.arange(cols) + ox) % 1024
    img = g[np.ix_(yy, xx)] + rng.uniform(-1.5, 1.5, (rows, cols))
    img = np.clip(img, 0, 255)
    if integer:
        return np.rint(img).astype(np.uint8)
    return img.astype(np.float32)


def normal_w(rows, cols, seed=1234):
    return np.random.default_rng(seed).standard_normal((rows, cols)).astype(np.float32)


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
