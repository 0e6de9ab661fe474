"""GPU parity at the scale of BASELINE.json's configs (reduced point counts, the configs' own densities and
parameters): the whole path -- shift, kNN rows, normals, plane growing (parallel engine), labels, raster --
against the CPU oracle, bit for bit, plus the size-independent properties the domain offers (row order,
self first, unit normals, canonical plane ids, both grower engines agree)."""
import numpy as np
import pytest

import oracle_lib as O
from buildingsegment_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from buildingsegment_b200 import lib

    c = lib.Context(0)
    yield c
    c.close()


def _crop(xyz, target):
    """Spatial crop at the full cloud's density (the corner square holding ~target points)."""
    if len(xyz) <= target:
        return xyz
    x = xyz[:, 0] - xyz[:, 0].min()
    y = xyz[:, 1] - xyz[:, 1].min()
    side = np.sqrt(target / len(xyz)) * max(x.max(), y.max())
    return np.ascontiguousarray(xyz[(x < side) & (y < side)])


CONFIGS = [
    # name, generator call, crop, parameters
    ("C1_building_1M", lambda: synth.make("C1", 1_000_000), None, dict()),
    ("C2_block_density_10M", lambda: synth.make("C2", 10_000_000), 1_500_000, dict()),
    ("C3_aerial_16bit", lambda: synth.make("C3", 4_000_000, size=280.0), None, dict(K=16, radius=1000.0)),
    ("C4_voxels", lambda: synth.make("C4", 1_500_000, bits=10), None,
     dict(K=16, radius=2.5, max_nn=50, th_thickness=3, bin=4, bin_height=16)),
]


@pytest.mark.parametrize("name,gen,crop,kw", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_config_matches_oracle(ctx, name, gen, crop, kw):
    from buildingsegment_b200 import lib

    xyz = gen()
    if crop:
        xyz = _crop(xyz, crop)
    n = len(xyz)
    p = lib.default_params(**kw)
    P = O.pipeline(xyz, **kw)
    mn, mx, xs = ctx.set_points(xyz)
    assert np.array_equal(xs, P["xyz"])
    neigh, nrm, _ = ctx.knn_normals(p)
    bad = np.nonzero((neigh != P["neigh"]).any(1))[0]
    assert len(bad) == 0, (name, len(bad), bad[:5])
    assert np.array_equal(nrm.view(np.int64), P["normals"].view(np.int64))
    # properties: self (or an exact duplicate with a lower index) first, distances ascending, ties by index
    rows = neigh.astype(np.int64)
    valid = rows >= 0
    d = xs[np.where(valid, rows, 0)].astype(np.int64) - xs[:, None, :].astype(np.int64)
    d2 = np.where(valid, (d * d).sum(2), np.iinfo(np.int64).max)
    assert np.all(d2[:, 0] == 0)
    assert np.all(np.diff(d2, axis=1) >= 0)
    tie = (np.diff(d2, axis=1) == 0) & valid[:, 1:]
    assert np.all(np.diff(rows, axis=1)[tie] > 0)
    nn = np.sqrt((nrm * nrm).sum(1))
    assert np.all(np.abs(nn - 1.0) < 1e-12) and np.all(nrm[:, 2] >= 0)
    # plane growing, parallel engine
    g = P["grow"]
    pidx, label, npl = ctx.grow_planes(lib.default_params(grow_mode=0, **kw))
    assert npl == g.n_planes
    assert np.array_equal(pidx, g.plane_idx) and np.array_equal(label, g.label)
    seeds, normals, centers, off, idx = ctx.get_planes(npl)
    assert np.array_equal(seeds, g.plane_seed) and np.all(np.diff(seeds) > 0)  # canonical ids: seed order
    assert np.array_equal(off, g.plane_off) and np.array_equal(idx, g.point_idx)
    assert np.array_equal(normals.view(np.int64), g.plane_normal.view(np.int64))
    assert np.array_equal(centers, g.plane_center)
    assert label.max(initial=0) <= npl and label.min(initial=0) >= 0
    # raster
    W, H = ctx.raster_size(p)
    img, a, b, c, th = ctx.raster(p)
    oimg = O.raster(xs, mx[2] - mn[2], W, H, bin=p.bin, bin_height=p.bin_height, bias=p.count_bias)
    assert np.array_equal(img.view(np.int64), oimg.view(np.int64))
    t = ctx.timings()
    print(name, n, "planes", npl, {k: round(t[k], 2) for k in ("sort", "cells", "knn", "knn_fallback", "grow", "raster")},
          {k: t[k] for k in ("n_unresolved", "n_big_cells", "grow_rounds")})


def test_both_engines_agree_at_5m(ctx):
    """No oracle at this size: the sequential engine (one warp, the reference's order) checks the parallel one."""
    from buildingsegment_b200 import lib

    xyz = _crop(synth.make("C2", 10_000_000), 5_000_000)
    ctx.set_points(xyz)
    ctx.knn_normals(lib.default_params(), want_neigh=False, want_normals=False)
    p0, l0, n0 = ctx.grow_planes(lib.default_params(grow_mode=0))
    p1, l1, n1 = ctx.grow_planes(lib.default_params(grow_mode=1))
    assert n0 == n1 and np.array_equal(p0, p1) and np.array_equal(l0, l1)
