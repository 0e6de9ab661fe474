"""CPU tests of the oracle: shared arithmetic vs numpy, kNN vs scipy, normals vs eigh, and the
region grower / raster restatements vs the reference's own lines (oracle/_ref)."""
import os
import tempfile

import numpy as np
import pytest

import cases
import oracle_lib as O


def test_shared_trig_and_log_match_libm():
    L = O.orc()
    xs = np.linspace(-1.0, 1.0, 4001)
    assert max(abs(L.orc_acos(float(v)) - np.arccos(v)) for v in xs) < 1e-15
    xs = np.linspace(0.0, 3.3, 4001)
    assert max(abs(L.orc_cos(float(v)) - np.cos(v)) for v in xs) < 5e-16
    xs = np.concatenate([np.linspace(1.0, 4.0, 2001), np.logspace(0, 15, 2001)])
    for v in xs[1:]:
        assert abs(L.orc_log(float(v)) - np.log(v)) <= 2 * np.spacing(np.log(v))
    assert L.orc_log(1.0) == 0.0
    assert L.orc_acos(1.0) == 0.0 and L.orc_acos(-1.0) == np.pi


def test_eigen_special_cases():
    L = O.orc()
    out = np.zeros(3)
    ev = np.zeros(3)
    L.orc_eigen(np.array([1.0, 0, 0, 1.0, 0, 1.0]), out, ev)  # identity -> +z (cov = I when < 3 nbrs)
    assert out.tolist() == [0.0, 0.0, 1.0]
    L.orc_eigen(np.zeros(6), out, ev)  # all-duplicate neighbourhood -> zero vector
    assert out.tolist() == [0.0, 0.0, 0.0]
    L.orc_eigen(np.array([0.5, 0, 0, 2.0, 0, 3.0]), out, ev)  # diagonal: smallest axis
    assert out.tolist() == [1.0, 0.0, 0.0]
    rng = np.random.default_rng(1)
    for _ in range(300):
        A = rng.normal(size=(3, 6)) * rng.uniform(0.01, 100, (3, 1))
        Cm = A @ A.T
        cov = np.array([Cm[0, 0], Cm[0, 1], Cm[0, 2], Cm[1, 1], Cm[1, 2], Cm[2, 2]])
        L.orc_eigen(cov, out, ev)
        w, v = np.linalg.eigh(Cm)
        assert np.allclose(np.sort(ev), w, rtol=1e-9, atol=1e-9 * w[-1])
        if w[1] - w[0] > 1e-6 * w[-1]:
            assert abs(abs(out @ v[:, 0]) - 1.0) < 1e-9


@pytest.mark.parametrize("case", ["building", "quantised", "sparse", "voxels"])
def test_knn_matches_ckdtree(case):
    from scipy.spatial import cKDTree

    xyz, *_ = O.bbox_shift(getattr(cases, case)())
    if len(xyz) > 30000:
        xyz = xyz[:30000]
    idx, d2 = O.knn(xyz, 50, cell=100)
    X = xyz.astype(np.float64)
    dd, ii = cKDTree(X).query(X, k=50)
    assert np.array_equal(np.round(dd ** 2).astype(np.int64), d2)
    # canonical order: (d2, idx) strictly increasing along each row
    key = d2.astype(np.float64) * 2.0 ** 32 + idx
    assert np.all(np.diff(key, axis=1) > 0)
    # the reported d2 are the true ones
    diff = X[idx[:, 7]] - X
    assert np.array_equal((diff ** 2).sum(1).astype(np.int64), d2[:, 7])


def test_knn_small_n_pads_with_minus_one():
    xyz = cases.tiny(9)
    idx, d2 = O.knn(xyz, 15, cell=100)
    assert np.all(idx[:, 9:] == -1) and np.all(idx[:, :9] >= 0)
    assert np.all(idx[:, 0] == np.arange(9))


@pytest.mark.parametrize("case", ["building", "block"])
def test_normals_match_eigh(case):
    P = O.pipeline(getattr(cases, case)(30000))
    X = P["xyz"].astype(np.float64)
    nh = P["n_hyb"]
    rng = np.random.default_rng(0)
    worst = 0.0
    for i in rng.choice(len(X), 2000, replace=False):
        k = nh[i]
        n = P["normals"][i]
        assert abs(np.linalg.norm(n) - 1.0) < 1e-12 and n[2] >= 0
        if k < 3:
            assert n.tolist() == [0.0, 0.0, 1.0]
            continue
        Q = X[P["knn"][i, :k]]
        w, v = np.linalg.eigh(np.cov(Q.T, bias=True))
        if w[1] - w[0] < 1e-3 * max(w[2], 1e-30):
            continue
        worst = max(worst, np.arccos(min(1.0, abs(v[:, 0] @ n))))
    assert worst < 1e-3  # north_star tolerance: 1e-3 rad


def _same_grow(a, b):
    assert a.n_planes == b.n_planes
    for f in ("plane_idx", "label", "plane_seed", "plane_off", "point_idx", "plane_center"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(a.plane_normal.view(np.int64), b.plane_normal.view(np.int64))


needs_ref = pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("case,kw", [
    ("building", dict(n=40000, order="shuffled")),
    ("building", dict(n=40000, order="scan")),
    ("block", dict(n=40000)),
    ("quantised", dict(n=20000)),
    ("far_offset", dict()),
    ("grid_plane", dict(nx=60, ny=60, order="shuffled")),
])
def test_grower_bit_exact_vs_reference_lines(case, kw):
    P = O.pipeline(getattr(cases, case)(**kw))
    r = O.ref_grow(P["xyz"], P["normals"], P["neigh"])
    _same_grow(P["grow"], r)
    # set_plane_color: same libc rand() sequence painted by the oracle
    rgb = O.libc_plane_colors(r.n_planes)
    assert np.array_equal(O.paint(len(P["xyz"]), r.plane_off, r.point_idx, rgb), r.colors)


@needs_ref
def test_grower_clutter_normals_order_dependence():
    """5 % clutter normals, row-major order => zero planes, thousands of orphans."""
    xyz = cases.grid_plane(100, 100, 30, order="row")
    P = O.pipeline(xyz)
    rng = np.random.default_rng(4)
    nrm = P["normals"].copy()
    bad = rng.random(len(nrm)) < 0.05
    nrm[bad] = [1.0, 0.0, 0.0]
    g = O.grow(P["xyz"], nrm, P["neigh"])
    r = O.ref_grow(P["xyz"], nrm, P["neigh"])
    _same_grow(g, r)
    assert ((g.plane_idx > 0) & (g.label == 0)).sum() > 100  # orphans exist


@needs_ref
def test_grower_int32_centre_overflow():
    """the int32 centroid wraps at 2^31 and the plane silently stops growing."""
    P = O.pipeline(cases.far_offset())
    g = P["grow"]
    r = O.ref_grow(P["xyz"], P["normals"], P["neigh"])
    _same_grow(g, r)
    assert g.n_planes >= 1 and (np.abs(g.plane_center.astype(np.int64)) > 10 ** 6).any()  # garbage centre
    assert int(g.plane_off[1]) < 10000  # stopped long before the 22 500 points of the plane


@needs_ref
@pytest.mark.parametrize("case,kw", [("building", dict(n=50000)), ("block", dict(n=50000)), ("tiny", dict(n=200)),
                                     ("count_sweep", dict())])
def test_raster_vs_reference_lines(case, kw):
    import cv2

    xyz = getattr(cases, case)(**kw)
    with tempfile.TemporaryDirectory() as d:
        xs, W, H, img_ref = O.ref_raster(xyz, d)
        x2, mn, mx, wh = O.bbox_shift(xyz)
        assert np.array_equal(xs, x2) and (W, H) == (int(wh[0]), int(wh[1]))
        img = O.raster(x2, mx[2] - mn[2], W, H)
        assert np.array_equal(img[..., 0], img_ref[..., 0])
        assert np.array_equal(img[..., 2], img_ref[..., 2])
        assert np.array_equal(img[..., 1].view(np.int64), img_ref[..., 1].view(np.int64))  # same libm log: bit-identical
        a, b, c, _ = O.save_image(img)
        for name, mine in (("平均高度.png", a), ("像素数量.png", b), ("像素数量+高度.png", c)):
            raw = np.frombuffer(open(os.path.join(os.fsencode(d), name.encode("gbk")), "rb").read(), np.uint8)
            png = cv2.imdecode(raw, cv2.IMREAD_COLOR)[..., ::-1]
            assert np.array_equal(png, mine), name
