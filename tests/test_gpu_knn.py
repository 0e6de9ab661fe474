"""GPU parity, stages a3-a5: bbox/shift, Morton binning primitives, exact kNN rows and PCA normals,
through the C ABI (buildingsegment_b200/lib.py -> libbseg.so), against the CPU oracle on the same
seeded inputs.  Bit-exact for indices AND normals (shared IEEE-only arithmetic); the north_star
tolerance (1e-4 rel / 1e-3 rad) is asserted too, against numpy eigh in test_oracle.py."""
import numpy as np
import pytest

import cases
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from buildingsegment_b200 import lib

    c = lib.Context(0)
    yield c
    c.close()


def test_exclusive_scan(ctx):
    rng = np.random.default_rng(0)
    for n in (1, 5, 4095, 4096, 4097, 1_000_003, 20_000_000):
        d = rng.integers(0, 5, n).astype(np.uint32)
        out = ctx.debug_exclusive_scan(d)
        ref = np.concatenate([[0], np.cumsum(d[:-1], dtype=np.uint64)]).astype(np.uint32)
        assert np.array_equal(out, ref), n


@pytest.mark.parametrize("n,bits", [(1, 8), (1000, 20), (3072, 63), (3073, 63), (500_000, 45), (3_000_000, 33)])
def test_radix_sort_pairs_stable(ctx, n, bits):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** min(bits, 62), n, dtype=np.uint64)
    if n > 10:
        keys[: n // 2] = keys[n // 2: n // 2 * 2]  # many duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    k2, v2 = ctx.debug_sort_pairs(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, vals[order])


def _check_case(ctx, xyz, **kw):
    from buildingsegment_b200 import lib

    p = lib.default_params(**kw)
    mn, mx, shifted = ctx.set_points(xyz)
    xs, omn, omx, _ = O.bbox_shift(xyz)
    assert np.array_equal(mn, omn) and np.array_equal(mx, omx)
    assert np.array_equal(shifted, xs)
    neigh, nrm, curv = ctx.knn_normals(p, want_curvature=True)
    kq = max(p.max_nn, p.K)
    idx, d2 = O.knn(xs, kq, cell=max(1, int(np.ceil(p.radius))))
    onrm, ocurv, nh = O.normals(xs, idx, d2, p.radius, p.max_nn)
    bad = np.nonzero((neigh != idx[:, : p.K]).any(1))[0]
    assert len(bad) == 0, (len(bad), bad[:5], neigh[bad[:2]], idx[bad[:2], : p.K])
    # normals: bit-exact against the shared-arithmetic oracle
    nb = np.nonzero((nrm.view(np.int64) != onrm.view(np.int64)).any(1))[0]
    assert len(nb) == 0, (len(nb), nb[:5], nrm[nb[:3]], onrm[nb[:3]], nh[nb[:3]])
    assert np.array_equal(curv.view(np.int64), ocurv.view(np.int64))
    # north_star tolerance, stated: 1e-4 relative / 1e-3 rad (trivially met by bit-exactness)
    assert np.all(np.abs(nrm - onrm) <= 1e-4 * np.maximum(np.abs(onrm), 1e-30))
    return ctx.timings()


@pytest.mark.parametrize("case,kw", [
    ("building", dict(n=60000)),
    ("building", dict(n=60000, order="scan")),
    ("block", dict(n=120000)),
    ("quantised", dict(n=50000)),
    ("voxels", dict(n=40000)),
    ("sparse", dict()),
    ("far_offset", dict()),
    ("grid_plane", dict(nx=80, ny=80, order="shuffled")),
])
def test_knn_and_normals_match_oracle(ctx, case, kw):
    t = _check_case(ctx, getattr(cases, case)(**kw))
    print(case, {k: t[k] for k in ("sort", "cells", "knn", "knn_fallback", "n_unresolved", "n_big_cells")})


@pytest.mark.parametrize("n", [1, 2, 3, 9, 14, 15, 16, 40])
def test_tiny_clouds(ctx, n):
    _check_case(ctx, cases.tiny(n))


def test_all_duplicates(ctx):
    xyz = np.tile(np.array([[5, 6, 7]], np.int32), (100, 1))
    _check_case(ctx, xyz)


@pytest.mark.parametrize("kw", [dict(K=16, radius=250.0, max_nn=30), dict(K=8, radius=40.0, max_nn=50),
                                dict(K=15, radius=100.0, cell=400), dict(K=32, radius=300.0, max_nn=64)])
def test_parameter_variants(ctx, kw):
    _check_case(ctx, cases.block(60000), **kw)


def test_voxel_units(ctx):
    """C4-style: coordinates in voxel units, radius 2.5 voxels => lattice ties everywhere."""
    _check_case(ctx, cases.voxels(40000), K=16, radius=2.5, max_nn=50)


def test_dense_groups_are_streamed_in_chunks(ctx):
    """30 000 points in a 300 mm cube: far more than the staging area of a group holds => the group kernel
    streams the block in chunks (and every radius is crowded: the max_nn nearest come from further passes)."""
    rng = np.random.default_rng(11)
    xyz = rng.integers(0, 300, (30000, 3)).astype(np.int32)
    _check_case(ctx, xyz, K=15, radius=100.0, max_nn=50, cell=100)
    _check_case(ctx, xyz, K=16, radius=100.0, max_nn=20, cell=100)
    # max_nn below K: the group kernel leaves the hybrid sets of crowded radii to the per-cell kernels
    _check_case(ctx, xyz, K=15, radius=100.0, max_nn=10, cell=100)


def test_max_nn_below_k(ctx):
    _check_case(ctx, cases.block(60000), K=15, radius=100.0, max_nn=10)
    _check_case(ctx, cases.building(40000), K=16, radius=150.0, max_nn=5)


def test_per_cell_kernels_alone(ctx, monkeypatch):
    """BSEG_KNN_GROUPS=0: the per-cell kernels (the path of K other than 15/16) serve the whole cloud, including
    the global-memory streaming variant for blocks of more than 512 candidates."""
    monkeypatch.setenv("BSEG_KNN_GROUPS", "0")
    rng = np.random.default_rng(11)
    xyz = rng.integers(0, 300, (30000, 3)).astype(np.int32)
    t = _check_case(ctx, xyz, K=15, radius=100.0, max_nn=50, cell=100)
    assert t["n_big_cells"] > 0
    _check_case(ctx, cases.block(60000))
    _check_case(ctx, cases.quantised(30000))


def test_extent_limit_is_loud(ctx):
    from buildingsegment_b200 import lib

    xyz = np.array([[0, 0, 0], [1 << 23, 0, 0]], np.int32)
    with pytest.raises(lib.BsegError):
        ctx.set_points(xyz)
