"""GPU parity of the round-2 ABI additions, through the C ABI against the CPU oracle:
  bseg_paint with a filtered / reordered plane list (my_function.cpp:260-275 paints what it is given),
  bseg_set_grow_offset (seg_plane on a cloud buildingSeg has not shifted: the int32 centroid wrap is in the caller's
  coordinates, my_function.cpp:190,227,242-250),
  un-normalised caller normals (exact model forced),
  bseg_knn_device_results + bseg_import_neigh_normals_device (rows / normals stay on the device, no kNN on import),
  bseg_halo_check / bseg_set_origin / bseg_set_owned / bseg_device_results (the multi-GPU entry points, one GPU)."""
import ctypes as C

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


def _dev_array(ptr, shape, typestr):
    import torch

    class _H:
        pass

    h = _H()
    h.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device="cuda:0")


def test_paint_subset_and_reordered(ctx):
    from buildingsegment_b200 import lib

    xyz = cases.block(120000)
    p = lib.default_params()
    ctx.set_points(xyz)
    ctx.knn_normals(p, want_neigh=False, want_normals=False)
    _, _, npl = ctx.grow_planes(p)
    g = O.pipeline(xyz)["grow"]
    assert npl == g.n_planes and npl >= 4
    rng = np.random.default_rng(3)
    for ids in (np.arange(npl, 0, -1), rng.permutation(npl)[: npl // 2] + 1, np.array([2, 2, 1]), np.zeros(0, np.int64)):
        ids = np.asarray(ids, np.int32)
        rgb = rng.integers(55, 255, (len(ids), 3)).astype(np.uint16)
        got = ctx.paint(rgb, plane_ids=ids)
        # the reference's loop: black, then every listed plane in order, later ones overwrite
        want = np.zeros((len(xyz), 3), np.uint16)
        for q, pid in enumerate(ids):
            want[g.point_idx[g.plane_off[pid - 1]: g.plane_off[pid]]] = rgb[q]
        assert np.array_equal(got, want)
    # all planes, canonical call == label path
    rgb = O.libc_plane_colors(npl)
    assert np.array_equal(ctx.paint(rgb), ctx.paint(rgb, plane_ids=np.arange(1, npl + 1)))
    # a colour count that disagrees with the plane count is refused, not read past the end
    with pytest.raises(lib.BsegError) as e:
        ctx.paint(rgb[:-1])
    assert e.value.code == -1
    with pytest.raises(lib.BsegError):
        ctx.paint(rgb[:1], plane_ids=np.array([npl + 1]))


@pytest.mark.parametrize("mode", [1, 0])
def test_grow_in_unshifted_coordinates(ctx, mode):
    """seg_plane handed the raw cloud: centroid sums wrap where the RAW coordinates make them wrap."""
    from buildingsegment_b200 import lib

    xyz = cases.far_offset() + np.array([700000, -250000, 40000], np.int32)
    p = lib.default_params(grow_mode=mode)
    mn, mx, xs = ctx.set_points(xyz)
    neigh, nrm, _ = ctx.knn_normals(p)
    g_raw = O.grow(xyz, nrm, neigh)       # the reference's arithmetic on the caller's coordinates
    g_shift = O.grow(xs, nrm, neigh)
    assert not np.array_equal(g_raw.plane_idx, g_shift.plane_idx)  # the case is sensitive to the coordinates
    ctx.set_grow_offset(mn)
    try:
        pidx, label, npl = ctx.grow_planes(p)
        assert npl == g_raw.n_planes and np.array_equal(pidx, g_raw.plane_idx) and np.array_equal(label, g_raw.label)
        seeds, normals, centers, off, idx = ctx.get_planes(npl)
        assert np.array_equal(centers, g_raw.plane_center) and np.array_equal(idx, g_raw.point_idx)
    finally:
        ctx.set_grow_offset(None)
    pidx, label, npl = ctx.grow_planes(p)
    assert npl == g_shift.n_planes and np.array_equal(pidx, g_shift.plane_idx)


@pytest.mark.parametrize("mode", [1, 0])
def test_grow_with_scaled_normals(ctx, mode):
    """Caller normals of magnitude 1e6 (seg_plane accepts any vectors): decisions must still be the reference's."""
    from buildingsegment_b200 import lib

    xyz = cases.building(40000)
    P = O.pipeline(xyz)
    rng = np.random.default_rng(12)
    nrm = P["normals"] * (10.0 ** rng.uniform(0, 6, (len(xyz), 1)))
    g = O.grow(P["xyz"], nrm, P["neigh"])
    p = lib.default_params(grow_mode=mode)
    ctx.set_points(xyz)
    ctx.knn_normals(p, want_neigh=False, want_normals=False)
    ctx.override(p, normals=nrm)
    pidx, label, npl = ctx.grow_planes(p)
    assert npl == g.n_planes and np.array_equal(pidx, g.plane_idx) and np.array_equal(label, g.label)
    seeds, normals, centers, off, idx = ctx.get_planes(npl)
    assert np.array_equal(normals.view(np.int64), g.plane_normal.view(np.int64))


@pytest.mark.parametrize("case,kw", [("block", dict(n=100000)), ("quantised", dict(n=40000)), ("tiny", dict(n=9))])
def test_device_export_import_roundtrip(case, kw):
    """Rows / normals leave one context as device arrays and enter another that never runs the kNN stage; the planes
    grown there are the oracle's."""
    import torch

    from buildingsegment_b200 import lib

    xyz = getattr(cases, case)(**kw)
    P = O.pipeline(xyz)
    p = lib.default_params()
    a, b = lib.Context(0), lib.Context(0)
    try:
        a.set_points(xyz)
        a.run_device(p, lib.RUN_KNN)
        d_neigh, d_nrm = a.knn_device_results(p)
        n = len(xyz)
        neigh = _dev_array(d_neigh, (n, p.K), "<i4")
        nrm = _dev_array(d_nrm, (n, 3), "<f8")
        assert np.array_equal(neigh.cpu().numpy(), P["neigh"])
        assert np.array_equal(nrm.cpu().numpy().view(np.int64), P["normals"].view(np.int64))
        neigh2, nrm2 = neigh.clone(), nrm.clone()
        torch.cuda.synchronize()
        b.set_points(xyz)
        b.reset_counters()
        b.import_neigh_normals_device(p, neigh2.data_ptr(), nrm2.data_ptr())
        pidx, label, npl = b.grow_planes(p)
        t = b.timings()
        assert t["knn"] == 0.0  # the kNN stage did not run in the importing context
        g = P["grow"]
        assert npl == g.n_planes and np.array_equal(pidx, g.plane_idx) and np.array_equal(label, g.label)
    finally:
        a.close()
        b.close()


def _halo_rule(xs, neigh, d2k, owned, x_lo, x_hi, halo, has_l, has_r):
    """numpy restatement of bseg_halo_check: owned points whose K-th neighbour may be farther than the nearest point
    the rank cannot see (halo + distance to a face that has a neighbour rank)."""
    x = xs[:, 0].astype(np.int64)
    reach = np.full(len(xs), np.iinfo(np.int64).max)
    if has_l:
        reach = np.minimum(reach, x - x_lo + halo)
    if has_r:
        reach = np.minimum(reach, x_hi - 1 - x + halo)
    short = neigh[:, -1] < 0
    bad = owned & (reach < np.iinfo(np.int64).max) & (short | (d2k > reach * reach))
    return int(bad.sum())


def test_halo_check_origin_owned_device_results(ctx):
    """One slab cut out of a cloud, on one GPU: shared origin, owned prefix, halo sufficiency count == the numpy rule,
    outer faces are not checked, device result pointers hold what the host exports return."""
    from buildingsegment_b200 import lib

    full = cases.block(150000)
    origin = full.min(axis=0).astype(np.int32)
    x = full[:, 0]
    x_lo, x_hi = int(origin[0]) + 12000, int(origin[0]) + 26000
    own = (x >= x_lo) & (x < x_hi)
    n_owned = int(own.sum())
    p = lib.default_params()
    ctx.set_origin(origin)
    try:
        halo = 200
        while True:  # as slabs.segment_tile does: double the halo until the check passes (sparse clutter needs metres)
            hal = ~own & (x >= x_lo - halo) & (x < x_hi + halo)
            idx_full = np.nonzero(own | hal)[0]  # global-index order: the tie order of the rows is the cloud's
            local = np.ascontiguousarray(full[idx_full])
            mn, mx, xs = ctx.set_points(local)
            assert np.array_equal(mn, origin)                                   # the shared origin was subtracted ...
            assert np.array_equal(xs, local - origin[None, :])                  # ... from every point
            neigh, nrm, _ = ctx.knn_normals(p)
            last = neigh[:, -1]
            d = xs[np.clip(last, 0, len(xs) - 1)].astype(np.int64) - xs.astype(np.int64)
            d2k = (d * d).sum(axis=1)
            owned_mask = own[idx_full]
            lo_s, hi_s = x_lo - int(origin[0]), x_hi - int(origin[0])
            for h in (50, 150, halo, 10 * halo):
                got = ctx.halo_check(lo_s, hi_s, h)
                assert got == _halo_rule(xs, neigh, d2k, owned_mask, lo_s, hi_s, h, True, True), h
                # outer faces (no neighbour rank): only the other face is checked; both outer = nothing to check
                assert ctx.halo_check(-2**31, hi_s, h) == _halo_rule(xs, neigh, d2k, owned_mask, lo_s, hi_s, h, False, True)
                assert ctx.halo_check(lo_s, 2**31 - 1, h) == _halo_rule(xs, neigh, d2k, owned_mask, lo_s, hi_s, h, True, False)
                assert ctx.halo_check(-2**31, 2**31 - 1, h) == 0
            # bseg_set_owned: indices at or above n_owned are halo copies whatever their position
            k = len(local) // 3
            ctx.set_owned(k)
            assert ctx.halo_check(lo_s, hi_s, 50) == _halo_rule(xs, neigh, d2k, owned_mask & (np.arange(len(local)) < k),
                                                                lo_s, hi_s, 50, True, True)
            ctx.set_owned(len(local))
            if ctx.halo_check(lo_s, hi_s, halo) == 0:
                break
            assert halo < 50000
            halo *= 2
        assert halo > 200  # the first width was NOT sufficient: the loop was exercised
        # with the halo sufficient, owned rows / normals are the undivided cloud's
        Pf = O.pipeline(full)
        assert np.array_equal(idx_full[neigh[owned_mask]], Pf["neigh"][own])
        assert np.array_equal(nrm[owned_mask].view(np.int64), Pf["normals"][own].view(np.int64))
        # device result pointers
        pidx, label, npl = ctx.grow_planes(p)
        d_label, d_pidx, d_xyz = ctx.device_results()
        assert np.array_equal(_dev_array(d_label, (len(local),), "<i4").cpu().numpy(), label)
        assert np.array_equal(_dev_array(d_pidx, (len(local),), "<i4").cpu().numpy(), pidx)
        assert np.array_equal(_dev_array(d_xyz, (len(local), 3), "<i4").cpu().numpy(), xs)
        with pytest.raises(lib.BsegError):
            ctx.set_owned(len(local) + 1)
        # an origin above the cloud's minimum is refused
        ctx.set_origin(origin + 1)
        with pytest.raises(lib.BsegError):
            ctx.set_points(local)
    finally:
        ctx.set_origin(None)


def test_kernel_attributes_on_a_second_context():
    """Per-context (per-device) kernel attributes: a second context created after the first one ran still launches
    the sweeper / group kernels (they need the opt-in shared memory size)."""
    from buildingsegment_b200 import lib

    xyz = cases.building(30000)
    P = O.pipeline(xyz)
    p = lib.default_params()
    for _ in range(2):
        c = lib.Context(0)
        c.set_points(xyz)
        c.knn_normals(p, want_neigh=False, want_normals=False)
        _, label, npl = c.grow_planes(p)
        assert npl == P["grow"].n_planes and np.array_equal(label, P["grow"].label)
        c.close()


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("case,kw,r", [("block", dict(n=150000), 260.0), ("building", dict(n=80000), 150.0)])
def test_radius_search_growing(ctx, case, kw, r, mode):
    """BASELINE config C2's variant (bseg_params.grow_radius): the grower's neighbour list is the part of the K-row
    inside the radius; the exported rows stay the full kNN rows."""
    from buildingsegment_b200 import lib

    xyz = getattr(cases, case)(**kw)
    P = O.pipeline(xyz, grow_radius=r)
    P0 = O.pipeline(xyz)
    assert not np.array_equal(P["grow"].plane_idx, P0["grow"].plane_idx)  # the variant changes the result here
    p = lib.default_params(grow_mode=mode, grow_radius=r)
    ctx.set_points(xyz)
    neigh, nrm, _ = ctx.knn_normals(p)
    assert np.array_equal(neigh, P0["neigh"])
    g = P["grow"]
    pidx, label, npl = ctx.grow_planes(p)
    assert npl == g.n_planes and np.array_equal(pidx, g.plane_idx) and np.array_equal(label, g.label)
    seeds, normals, centers, off, idx = ctx.get_planes(npl)
    assert np.array_equal(idx, g.point_idx) and np.array_equal(normals.view(np.int64), g.plane_normal.view(np.int64))
    # and back to the reference's rows with the radius off
    pidx, label, npl = ctx.grow_planes(lib.default_params(grow_mode=mode))
    assert npl == P0["grow"].n_planes and np.array_equal(pidx, P0["grow"].plane_idx)


@pytest.mark.parametrize("case,kw", [("building", dict(n=60000)), ("block", dict(n=150000))])
def test_plane_classes(ctx, case, kw):
    """DetectedPlane equations and roof / facade / ground classes (my_function.h:41-46; bseg_plane_classes) against the
    numpy restatement, per plane and per point; argument checks."""
    from buildingsegment_b200 import lib

    xyz = getattr(cases, case)(**kw)
    p = lib.default_params()
    ctx.set_points(xyz)
    ctx.knn_normals(p, want_neigh=False, want_normals=False)
    pidx, label, npl = ctx.grow_planes(p)
    assert npl > 0
    seeds, normals, centers, off, idx = ctx.get_planes(npl)
    _, _, _, _, th = ctx.raster(p, want_image=False, want_png=False)
    eq, pc, pt = ctx.plane_classes(th)
    oeq, opc, opt = O.plane_classes(normals, centers, label, th)
    assert np.array_equal(eq.view(np.int64), oeq.view(np.int64))
    assert np.array_equal(pc, opc) and np.array_equal(pt, opt)
    assert set(np.unique(pc)) <= {1, 2, 3, 4} and (pt[label == 0] == 0).all()
    assert (pc == lib.CLASS_ROOF).any() or (pc == lib.CLASS_GROUND).any()
    with pytest.raises(lib.BsegError):
        ctx.plane_classes(th, facade_max_nz=0.8, roof_min_nz=0.5)
