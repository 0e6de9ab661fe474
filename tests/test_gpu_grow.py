"""GPU parity, stages a7-a10 and a13-a15: plane region growing (both engines), plane export, paint
and the raster, through the C ABI against the CPU oracle (which is pinned bit-for-bit to the
reference's own lines in test_oracle.py).  Everything here is bit-exact: planeIdx, labels,
pointIdx lists, plane models, colours, image doubles and PNG bytes."""
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


def _grow_and_compare(ctx, p, g, n):
    pidx, label, npl = ctx.grow_planes(p)
    assert npl == g.n_planes, (npl, g.n_planes)
    bad = np.nonzero(pidx != g.plane_idx)[0]
    assert len(bad) == 0, (len(bad), bad[:10], pidx[bad[:10]], g.plane_idx[bad[:10]])
    assert np.array_equal(label, g.label)
    seeds, normals, centers, off, idx = ctx.get_planes(npl)
    assert np.array_equal(seeds, g.plane_seed)
    assert np.array_equal(off, g.plane_off)
    assert np.array_equal(idx, g.point_idx)
    assert np.array_equal(centers, g.plane_center)
    assert np.array_equal(normals.view(np.int64), g.plane_normal.view(np.int64))
    rgb = O.libc_plane_colors(npl)
    colors = ctx.paint(rgb)
    assert np.array_equal(colors, O.paint(n, g.plane_off, g.point_idx, rgb))
    return ctx.timings()


CASES = [
    ("building", dict(n=60000, order="shuffled")),
    ("building", dict(n=60000, order="scan")),
    ("block", dict(n=120000)),
    ("quantised", dict(n=50000)),
    ("voxels", dict(n=40000)),
    ("far_offset", dict()),
    ("grid_plane", dict(nx=100, ny=100, order="shuffled")),
    ("grid_plane", dict(nx=100, ny=100, order="row")),
    ("sparse", dict()),
    ("tiny", dict(n=9)),
    ("tiny", dict(n=40)),
]


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("case,kw", CASES)
def test_grow_matches_oracle(ctx, case, kw, mode):
    from buildingsegment_b200 import lib

    xyz = getattr(cases, case)(**kw)
    p = lib.default_params(grow_mode=mode)
    ctx.set_points(xyz)
    ctx.knn_normals(p, want_neigh=False, want_normals=False)
    P = O.pipeline(xyz)
    t = _grow_and_compare(ctx, p, P["grow"], len(xyz))
    print(case, mode, "planes", P["grow"].n_planes, {k: t[k] for k in ("grow", "finalize", "grow_steps", "grow_rounds")})


@pytest.mark.parametrize("mode", [1, 0])
def test_grow_with_clutter_normals_and_orphans(ctx, mode):
    """clutter normals in row-major order => orphan cascade; drives the grower with
    caller-supplied normals through bseg_override_neigh_normals."""
    from buildingsegment_b200 import lib

    xyz = cases.grid_plane(100, 100, 30, order="row")
    P = O.pipeline(xyz)
    rng = np.random.default_rng(4)
    nrm = P["normals"].copy()
    nrm[rng.random(len(nrm)) < 0.05] = [1.0, 0.0, 0.0]
    g = O.grow(P["xyz"], nrm, P["neigh"])
    assert ((g.plane_idx > 0) & (g.label == 0)).sum() > 100
    p = lib.default_params(grow_mode=mode)
    ctx.set_points(xyz)
    ctx.knn_normals(p, want_neigh=False, want_normals=False)
    ctx.override(p, normals=nrm)
    _grow_and_compare(ctx, p, g, len(xyz))


@pytest.mark.parametrize("mode", [1, 0])
def test_grow_with_adversarial_rows(ctx, mode):
    """Rows with repeated ids, self not in slot 0 and a shuffled neighbour order (Q7)."""
    from buildingsegment_b200 import lib

    xyz = cases.building(30000)
    P = O.pipeline(xyz)
    rng = np.random.default_rng(9)
    neigh = P["neigh"].copy()
    rows = rng.choice(len(neigh), 3000, replace=False)
    neigh[rows, 3] = neigh[rows, 2]                       # duplicate id inside a row
    rows = rng.choice(len(neigh), 3000, replace=False)
    neigh[rows, 0], neigh[rows, 5] = neigh[rows, 5].copy(), neigh[rows, 0].copy()  # self moved to slot 5
    g = O.grow(P["xyz"], P["normals"], neigh)
    p = lib.default_params(grow_mode=mode)
    ctx.set_points(xyz)
    ctx.knn_normals(p, want_neigh=False, want_normals=False)
    ctx.override(p, neigh=neigh)
    _grow_and_compare(ctx, p, g, len(xyz))


@pytest.mark.parametrize("mode", [1, 0])
def test_grow_threshold_variants(ctx, mode):
    from buildingsegment_b200 import lib

    xyz = cases.block(80000)
    kw = dict(K=12, th_thickness=150, th_point_count=100, th_dot=0.95)
    p = lib.default_params(grow_mode=mode, **kw)
    ctx.set_points(xyz)
    ctx.knn_normals(p, want_neigh=False, want_normals=False)
    P = O.pipeline(xyz, **kw)
    _grow_and_compare(ctx, p, P["grow"], len(xyz))


@pytest.mark.parametrize("case,kw", [("building", dict(n=60000)), ("block", dict(n=120000)), ("tiny", dict(n=200)),
                                     ("voxels", dict(n=40000)), ("sparse", dict()), ("count_sweep", dict())])
def test_raster_matches_oracle(ctx, case, kw):
    from buildingsegment_b200 import lib

    xyz = getattr(cases, case)(**kw)
    p = lib.default_params()
    mn, mx, xs = ctx.set_points(xyz)
    W, H = ctx.raster_size(p)
    _, _, _, wh = O.bbox_shift(xyz)
    assert (W, H) == (int(wh[0]), int(wh[1]))
    img, a, b, c, th = ctx.raster(p)
    oimg = O.raster(xs, mx[2] - mn[2], W, H)
    assert th == O.orc().orc_ground_th(xs, len(xs), int(mx[2] - mn[2]), 1000)
    assert np.array_equal(img.view(np.int64), oimg.view(np.int64))
    oa, ob, oc, _ = O.save_image(oimg)
    assert np.array_equal(a, oa) and np.array_equal(b, ob) and np.array_equal(c, oc)


def test_raster_device_with_given_threshold(ctx):
    """bseg_raster_device: the image stays on the device and the ground threshold is the caller's (slabs of a tile)."""
    import torch

    from buildingsegment_b200 import lib

    xyz = cases.block(80000)
    p = lib.default_params()
    mn, mx, xs = ctx.set_points(xyz)
    for th in (None, 0.0, 3000.0):
        d_img, W, H = ctx.raster_device(p, th)

        class _H:
            pass

        h = _H()
        h.__cuda_array_interface__ = {"shape": (H, W, 3), "typestr": "<f8", "data": (d_img, False), "version": 2}
        img = torch.as_tensor(h, device="cuda:0").cpu().numpy()
        # channel 1 leaves the device as the exact weight sums; the log is the host's (bseg_count_channel)
        ch1 = np.ascontiguousarray(img[..., 1])
        lib.count_channel(ch1, p.count_bias)
        img[..., 1] = ch1
        want = O.raster(xs, mx[2] - mn[2], W, H) if th is None else O.raster_th(xs, th, W, H)
        assert np.array_equal(img.view(np.int64), want.view(np.int64)), th


def test_raster_bin_variants(ctx):
    from buildingsegment_b200 import lib

    xyz = cases.block(60000)
    p = lib.default_params(bin=250, bin_height=500, count_bias=3.0)
    mn, mx, xs = ctx.set_points(xyz)
    img, a, b, c, th = ctx.raster(p)
    H, W, _ = img.shape
    oimg = O.raster(xs, mx[2] - mn[2], W, H, bin=250, bin_height=500, bias=3.0)
    assert np.array_equal(img.view(np.int64), oimg.view(np.int64))


def test_segment_host_end_to_end(ctx):
    """bseg_segment_host == the staged calls == the oracle pipeline."""
    from buildingsegment_b200 import lib

    xyz = cases.block(100000)
    p = lib.default_params()
    P = O.pipeline(xyz)
    shifted = np.empty_like(xyz)
    label = np.empty(len(xyz), np.int32)
    W, H = int(P["wh"][0]), int(P["wh"][1])
    a = np.empty((H, W, 3), np.uint8)
    b = np.empty((H, W, 3), np.uint8)
    npl, W2, H2 = ctx.segment_host(p, xyz, shifted, label, a, b)
    assert (W2, H2) == (W, H) and npl == P["grow"].n_planes
    assert np.array_equal(shifted, P["xyz"]) and np.array_equal(label, P["grow"].label)
    oimg = O.raster(P["xyz"], P["mx"][2] - P["mn"][2], W, H)
    oa, ob, _, _ = O.save_image(oimg)
    assert np.array_equal(a, oa) and np.array_equal(b, ob)


@pytest.mark.parametrize("case,kw", [("building", dict(n=60000)), ("block", dict(n=120000)), ("voxels", dict(n=40000))])
def test_label_raster(ctx, case, kw):
    """bseg_label_raster: per pixel the label of the highest point (ties: lower index) and its plane colour."""
    from buildingsegment_b200 import lib

    xyz = getattr(cases, case)(**kw)
    p = lib.default_params()
    mn, mx, xs = ctx.set_points(xyz)
    ctx.knn_normals(p, want_neigh=False, want_normals=False)
    pidx, label, npl = ctx.grow_planes(p)
    rgb = O.libc_plane_colors(npl)
    W, H = ctx.raster_size(p)
    lab, img = ctx.label_raster(p, rgb)
    olab, oimg = O.label_raster(xs, label, W, H, rgb)
    assert np.array_equal(lab, olab) and np.array_equal(img, oimg)
    assert lab.max(initial=0) <= npl
    if npl:
        assert (lab > 0).any()


def test_stage_order_errors_are_loud(ctx):
    from buildingsegment_b200 import lib

    p = lib.default_params()
    ctx.set_points(cases.tiny(50))
    with pytest.raises(lib.BsegError):
        ctx.grow_planes(p)
    with pytest.raises(lib.BsegError):
        ctx.label_raster(p)
    with pytest.raises(lib.BsegError):
        lib.Context(99)
