"""Small seeded inputs shared by the oracle tests, the golden-vector generator and the GPU parity
tests.  Every case returns an UNSHIFTED int32 [N,3] cloud (millimetres)."""
import numpy as np

from buildingsegment_b200 import synth


def grid_plane(nx=150, ny=150, step=30, tilt=0.0, offset=(0, 0, 0), order="row", seed=0, jitter=0):
    """Planar lattice z = tilt*x (SURVEY A.4 / A.2-Q6 probes)."""
    rng = np.random.default_rng(seed)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    x = ix.ravel() * step
    y = iy.ravel() * step
    z = np.trunc(tilt * x).astype(np.int64)
    p = np.stack([x, y, z], 1).astype(np.int64)
    if jitter:
        p = p + rng.integers(-jitter, jitter + 1, p.shape)
    p = p + np.asarray(offset, np.int64)[None, :]
    if order == "shuffled":
        p = p[rng.permutation(len(p))]
    return np.ascontiguousarray(p, np.int32)


def building(n=60000, order="shuffled", seed=1001):
    return synth.single_building(n, seed=seed, order=order)


def block(n=120000, seed=1002):
    """A 40 x 40 m corner of a C2-style block: buildings + ground + vegetation clutter."""
    rng = np.random.default_rng(seed)
    pts = synth._block(rng, n, 0.0, 0.0, 40.0, 3, 0.15, "shuffled")
    return synth.to_mm(pts)


def quantised(n=50000, seed=1003):
    """C3-style: coarse lattice => exact distance ties and duplicate points."""
    rng = np.random.default_rng(seed)
    pts = synth._block(rng, n, 0.0, 0.0, 30.0, 2, 0.1, "generation")
    q = np.round(pts / 0.04) * 0.04
    return synth.to_mm(q)


def voxels(n=40000, seed=1004):
    return synth.voxel_scan(n, seed=seed, bits=7)


def sparse(n=3000, seed=7):
    """Very sparse volume: nearly every kNN needs ring expansion, most points have < 3 hybrid nbrs."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 20000, (n, 3)).astype(np.int32)


def tiny(n, seed=3):
    rng = np.random.default_rng(seed)
    return rng.integers(-500, 500, (n, 3)).astype(np.int32)


def far_offset(n=None, seed=5):
    """A plane 300 m from the cloud minimum (one stray point pins the bbox): the sum of x passes
    2^31 after ~7000 points and the int32 centroid wraps."""
    p = grid_plane(150, 150, 30, tilt=0.01, offset=(300000, 0, 0), order="row", seed=seed)
    return np.ascontiguousarray(np.concatenate([np.zeros((1, 3), np.int32), p], 0))


def count_sweep(kmax=1200, seed=11):
    """Pixel k of a row holds exactly k points at its corner (bilinear weight 1.0 on one tap), k = 1..kmax, plus a
    second row with fractional weights: the count channel log(count + 1) + 20 then sweeps densely through the uint8
    truncation boundaries of save_image's (uint8)(255.0 * v / max) (TMC3.cpp:159-164, 101-108)."""
    rng = np.random.default_rng(seed)
    xs, ys, zs = [], [], []
    for k in range(1, kmax + 1):
        xs.append(np.full(k, k * 100, np.int64))
        ys.append(np.zeros(k, np.int64))
        zs.append(rng.integers(5000, 9000, k))
        m = max(1, k // 3)
        xs.append(k * 100 + rng.integers(0, 100, m))
        ys.append(300 + rng.integers(0, 100, m))
        zs.append(rng.integers(5000, 9000, m))
    p = np.stack([np.concatenate(xs), np.concatenate(ys), np.concatenate(zs)], 1)
    p = p[rng.permutation(len(p))]
    return np.ascontiguousarray(p, np.int32)
