"""Synthetic building clouds for the BASELINE.json configs.

The reference ships no data (readme.txt:12 only names the author's local files), so every
workload is generated: float32 metres, as a PLY would hold them, converted to the int32
millimetres `ply::read(..., scale=1000)` produces (/root/reference/tmc3/ply.cpp:407-409:
``int32(double(float) * 1000.0)``, truncation toward zero).

Generators are numpy-only, seeded (PCG64) and deterministic.  They return ``xyz`` as a C-contiguous
``int32 [N,3]`` array (not yet shifted to the origin -- the buildingSeg constructor does that).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "to_mm",
    "single_building",
    "suburban_block",
    "aerial_tile",
    "voxel_scan",
    "city_tile",
    "make",
    "CONFIGS",
]


def to_mm(xyz_m: np.ndarray, scale: float = 1000.0) -> np.ndarray:
    """ply::read position conversion: float32 -> double * scale -> int32 truncation."""
    return np.trunc(xyz_m.astype(np.float32).astype(np.float64) * scale).astype(np.int32)


# ------------------------------------------------------------------------------------------------
# primitives: every surface is a parallelogram or triangle patch (origin, edge u, edge v)
# ------------------------------------------------------------------------------------------------
def _patch_points(rng, n, o, u, v, tri, sigma):
    a = rng.random(n)
    b = rng.random(n)
    if tri:
        flip = a + b > 1.0
        a = np.where(flip, 1.0 - a, a)
        b = np.where(flip, 1.0 - b, b)
    p = o[None, :] + a[:, None] * u[None, :] + b[:, None] * v[None, :]
    nrm = np.cross(u, v)
    nrm = nrm / np.linalg.norm(nrm)
    return p + rng.normal(0.0, sigma, n)[:, None] * nrm[None, :]


def _building_patches(cx, cy, z0, L, Wd, eave, pitch_deg, yaw_deg):
    """4 walls + 2 gable triangles + 2 roof slopes of a gable-roof box; returns (o,u,v,tri,area)."""
    c, s = np.cos(np.radians(yaw_deg)), np.sin(np.radians(yaw_deg))
    ex = np.array([c, s, 0.0])
    ey = np.array([-s, c, 0.0])
    ez = np.array([0.0, 0.0, 1.0])
    base = np.array([cx, cy, z0]) - 0.5 * L * ex - 0.5 * Wd * ey
    rise = 0.5 * Wd * np.tan(np.radians(pitch_deg))
    P = []
    # walls
    P.append((base, L * ex, eave * ez, False))
    P.append((base + Wd * ey, L * ex, eave * ez, False))
    P.append((base, Wd * ey, eave * ez, False))
    P.append((base + L * ex, Wd * ey, eave * ez, False))
    # gable triangles (ridge runs along ex)
    top0 = base + eave * ez
    P.append((top0, Wd * ey, 0.5 * Wd * ey + rise * ez, True))
    P.append((top0 + L * ex, Wd * ey, 0.5 * Wd * ey + rise * ez, True))
    # roof slopes
    P.append((top0, L * ex, 0.5 * Wd * ey + rise * ez, False))
    P.append((top0 + Wd * ey, L * ex, -0.5 * Wd * ey + rise * ez, False))
    out = []
    for o, u, v, tri in P:
        area = np.linalg.norm(np.cross(u, v)) * (0.5 if tri else 1.0)
        out.append((o, u, v, tri, area))
    return out


def _sample_patches(rng, patches, n, sigma):
    areas = np.array([p[4] for p in patches])
    counts = rng.multinomial(n, areas / areas.sum())
    chunks = [
        _patch_points(rng, int(k), o, u, v, tri, sigma)
        for (o, u, v, tri, _), k in zip(patches, counts)
        if k > 0
    ]
    return np.concatenate(chunks, axis=0) if chunks else np.zeros((0, 3))


def _ground(rng, n, x0, x1, y0, y1, amp, wavelen, sigma):
    x = rng.uniform(x0, x1, n)
    y = rng.uniform(y0, y1, n)
    z = amp * np.sin(2 * np.pi * x / wavelen) * np.cos(2 * np.pi * y / (1.3 * wavelen))
    return np.stack([x, y, z + rng.normal(0.0, sigma, n)], axis=1)


def _vegetation(rng, n, x0, x1, y0, y1, n_blobs):
    if n <= 0:
        return np.zeros((0, 3))
    cen = np.stack([rng.uniform(x0, x1, n_blobs), rng.uniform(y0, y1, n_blobs), rng.uniform(2.0, 6.0, n_blobs)], 1)
    rad = np.stack([rng.uniform(1.0, 3.0, n_blobs), rng.uniform(1.0, 3.0, n_blobs), rng.uniform(1.5, 4.0, n_blobs)], 1)
    which = rng.integers(0, n_blobs, n)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = rng.random(n) ** (1.0 / 3.0)
    return cen[which] + d * r[:, None] * rad[which]


def _order(rng, pts, order):
    if order == "shuffled":
        return pts[rng.permutation(len(pts))]
    if order == "scan":  # row order: y-major strips of 5 cm, then x
        key = np.lexsort((pts[:, 0], np.floor(pts[:, 1] / 0.05)))
        return pts[key]
    if order == "generation":
        return pts
    raise ValueError(order)


# ------------------------------------------------------------------------------------------------
# configs
# ------------------------------------------------------------------------------------------------
def single_building(n=1_000_000, seed=1001, order="shuffled", sigma=0.010):
    """C1: one 20 x 12 m gable-roof building, eaves 6 m, pitch 30 deg, sigma = 10 mm."""
    rng = np.random.default_rng(seed)
    patches = _building_patches(10.0, 6.0, 0.0, 20.0, 12.0, 6.0, 30.0, 0.0)
    pts = _sample_patches(rng, patches, n, sigma)
    return to_mm(_order(rng, pts, order))


def _block(rng, n, x0, y0, size, n_buildings, veg_frac, order):
    n_veg = int(n * veg_frac)
    patches = []
    for _ in range(n_buildings):
        L = rng.uniform(8.0, 20.0)
        Wd = rng.uniform(8.0, min(L, 14.0))
        patches += _building_patches(
            x0 + rng.uniform(15.0, size - 15.0), y0 + rng.uniform(15.0, size - 15.0), 0.0,
            L, Wd, rng.uniform(3.0, 9.0), rng.uniform(15.0, 45.0), rng.uniform(0.0, 180.0))
    area_b = sum(p[4] for p in patches)
    area_g = size * size
    n_surf = n - n_veg
    n_b = int(n_surf * area_b / (area_b + area_g))
    sigma = rng.uniform(0.010, 0.020)
    pts = np.concatenate([
        _sample_patches(rng, patches, n_b, sigma),
        _ground(rng, n_surf - n_b, x0, x0 + size, y0, y0 + size, 0.4, 60.0, sigma),
        _vegetation(rng, n_veg, x0, x0 + size, y0, y0 + size, max(8, n_buildings)),
    ], axis=0)
    return _order(rng, pts, order)


def suburban_block(n=10_000_000, seed=1002, order="shuffled"):
    """C2: 200 x 200 m block, 40 buildings + undulating ground + 15 % vegetation clutter."""
    rng = np.random.default_rng(seed)
    return to_mm(_block(rng, n, 0.0, 0.0, 200.0, 40, 0.15, order))


def aerial_tile(n=50_000_000, seed=1003, size=1000.0):
    """C3: aerial-LiDAR-style tile, flight-line order, coordinates quantised to 16 bit."""
    rng = np.random.default_rng(seed)
    nb = max(4, int(400 * (size / 1000.0) ** 2))
    pts = _block(rng, n, 0.0, 0.0, size, nb, 0.10, "generation")
    # nadir-style: thin out facades (|n_z| small is not known per point here: keep it simple and
    # sort into flight lines of 50 m swaths along x)
    swath = np.floor(pts[:, 1] / 50.0)
    pts = pts[np.lexsort((pts[:, 0], swath))]
    step = size / 65535.0
    q = np.round(pts / step) * step  # 16-bit lattice => exact ties and duplicates
    return to_mm(q)


def _voxel_block(seed, n, bits, ox, oy):
    """One scanned building voxelised to a 2^bits lattice: unique voxels of 3-voxel-thick shells (walls, roof, floor)."""
    rng = np.random.default_rng(seed)
    side = float(2 ** bits)
    patches = _building_patches(side / 2, side / 2, 0.1 * side, rng.uniform(0.6, 0.75) * side, rng.uniform(0.4, 0.5) * side,
                                rng.uniform(0.3, 0.4) * side, rng.uniform(20.0, 40.0), rng.uniform(0.0, 90.0))
    patches.append((np.array([0.0, 0.0, 0.1 * side]), np.array([side - 1, 0, 0.0]), np.array([0, side - 1, 0.0]), False,
                    side * side))
    cap = int(0.9 * 3 * sum(p[4] for p in patches))  # what 3-voxel shells of these surfaces can hold
    want = min(n, cap)
    keys = np.zeros(0, np.int64)
    for _ in range(12):
        need = want - len(keys)
        if need <= 0:
            break
        pts = _sample_patches(rng, patches, int(need * 1.7) + 1000, 1.0)
        v = np.clip(np.round(pts), 0, side - 1).astype(np.int64)
        k = (v[:, 2] << (2 * bits)) | (v[:, 1] << bits) | v[:, 0]  # z-major raster order
        keys = np.unique(np.concatenate([keys, k]))
    if len(keys) > want:
        keys = np.sort(rng.choice(keys, want, replace=False))
    mask = (1 << bits) - 1
    out = np.stack([(keys & mask) + ox, ((keys >> bits) & mask) + oy, keys >> (2 * bits)], axis=1)
    return np.ascontiguousarray(out, np.int32)


def voxel_scan(n=100_000_000, seed=1004, bits=10, workers=1):
    """C4: dense building scans voxelised to 2^bits integer lattices (unique voxels, 3-voxel shells), one scan per
    block, blocks side by side (block-major, raster order inside a block).  Coordinates are voxel units (scale 1)."""
    side = 2 ** bits
    per_block = int(0.9 * 3 * 2.2 * side * side)  # ~ what one block can hold (see _voxel_block)
    nb = max(1, -(-n // per_block))
    g = int(np.ceil(np.sqrt(nb)))
    counts = [n // nb + (1 if b < n % nb else 0) for b in range(nb)]
    jobs = [(seed + 1000 * b, counts[b], bits, (b % g) * side, (b // g) * side) for b in range(nb)]
    return np.concatenate(_map(_voxel_block_job, jobs, workers), axis=0)


def _voxel_block_job(a):
    return _voxel_block(*a)


def _map(fn, jobs, workers):
    """Blocks are independent (own seed each): generate them on `workers` processes.  fork()s -- call before CUDA is
    initialised in this process."""
    if workers <= 1 or len(jobs) <= 1:
        return [fn(j) for j in jobs]
    import multiprocessing as mp

    with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
        return pool.map(fn, jobs, chunksize=1)


def city_block(b, n_block, seed=1005, blocks_per_side=4, block=500.0):
    """Block b of the C5 tile (own seed: the tile does not depend on who generates which block)."""
    rng = np.random.default_rng(seed + 1000 * b)
    bx, by = b % blocks_per_side, b // blocks_per_side
    nbld = max(4, int(40 * (block / 200.0) ** 2))
    return to_mm(_block(rng, n_block, bx * block, by * block, block, nbld, 0.15, "shuffled"))


def _city_block_job(a):
    return city_block(*a)


def city_tile_counts(n, blocks_per_side=4):
    nb = blocks_per_side * blocks_per_side
    per = n // nb
    return [per if b < nb - 1 else n - per * (nb - 1) for b in range(nb)]


def city_tile(n=200_000_000, seed=1005, blocks_per_side=4, block=500.0, blocks=None, workers=1):
    """C5: city tile = blocks_per_side^2 suburban blocks, block-major, shuffled inside a block.  `blocks` selects a
    contiguous run of blocks (a rank's chunk of the tile: the tile's index order is block-major)."""
    counts = city_tile_counts(n, blocks_per_side)
    sel = range(len(counts)) if blocks is None else blocks
    jobs = [(b, counts[b], seed, blocks_per_side, block) for b in sel]
    return np.concatenate(_map(_city_block_job, jobs, workers), axis=0)


CONFIGS = {
    "C1": dict(fn=single_building, n=1_000_000),
    "C2": dict(fn=suburban_block, n=10_000_000),
    "C3": dict(fn=aerial_tile, n=50_000_000),
    "C4": dict(fn=voxel_scan, n=100_000_000),
    "C5": dict(fn=city_tile, n=200_000_000),
}


def make(name: str, n: int | None = None, **kw) -> np.ndarray:
    """Generate config `name` (C1..C5), optionally at a reduced point count `n`."""
    cfg = CONFIGS[name]
    xyz = cfg["fn"](n=cfg["n"] if n is None else n, **kw)
    return np.ascontiguousarray(xyz, dtype=np.int32)
