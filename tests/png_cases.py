"""Seeded byte images for the PNG writer tests (shared by the golden generator and the tests)."""
import numpy as np


def images():
    rng = np.random.default_rng(20261018)
    yield "noise_rgb", rng.integers(0, 256, (37, 53, 3)).astype(np.uint8)
    yield "zeros", np.zeros((64, 80, 3), np.uint8)
    g = np.zeros((120, 200, 3), np.uint8)
    g[..., 1] = np.add.outer(np.arange(120), np.arange(200)) % 256
    yield "gradient", g
    s = np.zeros((300, 400, 3), np.uint8)
    s[50:200, 100:300, 0] = rng.integers(200, 256, (150, 200))
    yield "sparse", s
    yield "grey", (rng.integers(0, 4, (90, 70)) * 60).astype(np.uint8)
    yield "grey_alpha", rng.integers(0, 256, (33, 21, 2)).astype(np.uint8)
    yield "rgba", rng.integers(0, 256, (20, 30, 4)).astype(np.uint8)
    yield "one_px", np.array([[[1, 2, 3]]], np.uint8)
    yield "row", rng.integers(0, 256, (1, 500, 3)).astype(np.uint8)
    yield "column", rng.integers(0, 256, (500, 1, 3)).astype(np.uint8)
    big = np.zeros((700, 900, 3), np.uint8)  # like image B of save_image: one channel, bright where occupied
    big[..., 1] = (rng.random((700, 900)) < 0.3) * rng.integers(230, 256, (700, 900))
    yield "raster_like", big
    rep = np.tile(rng.integers(0, 256, (4, 16, 3)).astype(np.uint8), (60, 40, 1))  # long matches, lazy matching
    yield "periodic", rep
    yield "incompressible_small", rng.integers(0, 256, (3, 3, 3)).astype(np.uint8)  # stored blocks win
