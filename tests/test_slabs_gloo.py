"""Multi-GPU tile logic (buildingsegment_b200/slabs.py) on the CPU: world_size 2 and 3 over gloo.

The compute backend here is a stand-in built on the CPU oracle (test infrastructure); what is under test is the host
logic of the N > 1 path: the partitioner, the halo exchange in global-index order, the halo sufficiency loop, the
assembly of the tile's rows / normals on the root, the label scatter and the stitched raster.  The bar is the product's:
  * labels and planeIdx of every chunk == those of the UNDIVIDED tile (seg_plane::get_planes over the whole cloud,
    my_function.cpp:180-258), bit for bit, plane count included;
  * kNN rows (as global indices, tie order included) and normals of every owned point == the undivided cloud's;
  * the stitched pixel columns == the raster of the undivided tile: doubles, ground threshold, save_image bytes.
The same path with the CUDA backend runs under torchrun on the GPU box (tests/test_gpu_tile.py).
"""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import oracle_lib as O


class OracleBackend:
    """Same contract as slabs.CudaBackend, computed by the oracle (CPU tests only)."""

    def __init__(self, K=15, radius=100.0, max_nn=50):
        self.p = types.SimpleNamespace(K=K, radius=radius, max_nn=max_nn, count_bias=20.0)

    @staticmethod
    def _shift(xyz, origin):
        return np.ascontiguousarray(xyz.numpy().astype(np.int64) - np.asarray(origin, np.int64)).astype(np.int32)

    def knn(self, local_xyz, origin, x_lo_s, x_hi_s, halo):
        p = self.p
        xs = self._shift(local_xyz, origin)
        idx, d2 = O.knn(xs, p.max_nn, cell=100)
        nrm, _, _ = O.normals(xs, idx, d2, p.radius, p.max_nn)
        neigh = np.ascontiguousarray(idx[:, : p.K])
        # bseg_halo_check: owned by position; a face at INT32_MIN / INT32_MAX has no neighbour rank
        x = xs[:, 0].astype(np.int64)
        has_l, has_r = x_lo_s != -2**31, x_hi_s != 2**31 - 1
        owned = np.ones(len(xs), bool)
        big = np.iinfo(np.int64).max
        reach = np.full(len(xs), big)
        if has_l:
            owned &= x >= x_lo_s
            reach = np.minimum(reach, x - x_lo_s + halo)
        if has_r:
            owned &= x < x_hi_s
            reach = np.minimum(reach, x_hi_s - 1 - x + halo)
        dk = d2[:, p.K - 1]
        bad = int(np.count_nonzero(owned & (reach < big) & ((dk < 0) | (dk > reach * reach))))
        if bad:
            return None, None, bad
        return torch.from_numpy(neigh), torch.from_numpy(nrm), 0

    def grow_tile(self, tile_xyz, neigh, normals):
        t = tile_xyz.numpy()
        xs = self._shift(tile_xyz, t.min(axis=0))
        g = O.grow(xs, np.ascontiguousarray(normals.numpy()), np.ascontiguousarray(neigh.numpy()))
        return torch.from_numpy(g.label.astype(np.int32)), torch.from_numpy(g.plane_idx.astype(np.int32)), int(g.n_planes)

    def raster(self, local_xyz, origin, ground_th):
        xs = self._shift(local_xyz, origin)
        W, H = int(xs[:, 0].max()) // 100 + 2, int(xs[:, 1].max()) // 100 + 2
        return torch.from_numpy(O.raster_th_sums(xs, ground_th, W, H))

    def count_channel(self, sums, bias):
        from buildingsegment_b200 import lib  # the product's host half (plain libm, no GPU needed)

        return lib.count_channel(sums, bias)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _chunks(n, world, skew):
    """Contiguous, deliberately unequal pieces of the tile (what reading stripes of a file gives)."""
    w = np.array([1.0 + skew * ((r * 7) % 3) for r in range(world)])
    cuts = np.concatenate([[0], np.round(np.cumsum(w) / w.sum() * n).astype(np.int64)])
    cuts[-1] = n
    return cuts


def _worker(rank, world, port, case, kw, halo, out_dir, want_raster):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from buildingsegment_b200 import slabs

        xyz = getattr(cases, case)(**kw)
        cuts = _chunks(len(xyz), world, 0.6)
        chunk = torch.from_numpy(np.ascontiguousarray(xyz[cuts[rank]: cuts[rank + 1]]))
        be = OracleBackend()
        r = slabs.segment_tile(be, chunk, halo=halo)
        P = r["partition"]
        out = dict(labels=r["labels"].numpy(), plane_idx=r["plane_idx"].numpy(), n_planes=r["n_planes"], halo=r["halo"],
                   n_halo=r["n_halo"], c0=cuts[rank], c1=cuts[rank + 1], owned_gid=r["owned_gid"].numpy(),
                   owned_rows=r["owned_rows"].numpy(), owned_normals=r["owned_normals"].numpy(), cuts=P.cuts,
                   owned_xyz=P.owned.numpy(), part_gid=P.gid.numpy(), origin=P.origin, owned_counts=P.owned_counts)
        if want_raster:
            img = slabs.raster_tile(be, chunk, r)
            out.update(image=img["image"].numpy(), a=img["png_a"].numpy(), b=img["png_b"].numpy(), x0=img["x0"], W=img["W"], H=img["H"],
                       th=img["ground_th"])
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), **out)
    finally:
        dist.destroy_process_group()


def _run(case, kw, halo, tmp_path, world=2, want_raster=False):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, kw, halo, str(tmp_path), want_raster), nprocs=world, join=True)
    return [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]


def _check_exact(res, xyz):
    P = O.pipeline(xyz)  # the undivided tile (shifted by its own minimum == the tile origin)
    g = P["grow"]
    seen = np.zeros(len(xyz), bool)
    for r in res:
        c0, c1 = int(r["c0"]), int(r["c1"])
        assert int(r["n_planes"]) == g.n_planes
        assert np.array_equal(r["labels"], g.label[c0:c1]), "labels of a chunk differ from the undivided tile"
        assert np.array_equal(r["plane_idx"], g.plane_idx[c0:c1]), "planeIdx of a chunk differs from the undivided tile"
        og = r["owned_gid"]
        assert np.all(np.diff(og) > 0)  # owned points in global-index order
        assert not seen[og].any()
        seen[og] = True
        assert np.array_equal(r["owned_rows"], P["neigh"][og]), "kNN rows of owned points differ from the undivided cloud"
        assert np.array_equal(r["owned_normals"].view(np.int64), P["normals"][og].view(np.int64)), "normals differ"
        assert np.array_equal(r["origin"], xyz.min(axis=0))
    assert seen.all()  # every point of the tile is owned by exactly one rank
    return P


@pytest.mark.parametrize("world,case,kw", [(2, "building", dict(n=30000, order="shuffled")),
                                           (3, "block", dict(n=45000)),
                                           (2, "quantised", dict(n=25000)),     # exact ties and duplicates across the face
                                           (3, "grid_plane", dict(nx=180, ny=50, order="shuffled"))])  # one plane over two faces
def test_tile_labels_are_the_undivided_tiles(world, case, kw, tmp_path):
    res = _run(case, kw, 400, tmp_path, world=world)
    xyz = getattr(cases, case)(**kw)
    P = _check_exact(res, xyz)
    assert P["grow"].n_planes > 0 and sum(int(r["n_halo"]) for r in res) > 0
    # the partitioner: slabs respect the cuts and hold about N / world points each
    cuts = res[0]["cuts"]
    for k, r in enumerate(res):
        x = r["owned_xyz"][:, 0]
        if k > 0:
            assert x.min() >= cuts[k - 1]
        if k < world - 1:
            assert x.max() < cuts[k]
        assert np.array_equal(r["owned_counts"], res[0]["owned_counts"])
    if case != "quantised":  # (a coarse lattice has few distinct x: the histogram cannot cut finer than a lattice plane)
        assert res[0]["owned_counts"].max() <= 1.25 * len(xyz) / world + 64


def test_insufficient_halo_is_doubled_until_it_suffices(tmp_path):
    """A halo narrower than some K-th neighbour distance: the sufficiency check fails on some rank, every rank
    doubles the halo and exchanges again; the result is still the undivided tile's."""
    kw = dict(n=30000, order="shuffled")
    res = _run("building", kw, 100, tmp_path)
    assert int(res[0]["halo"]) == int(res[1]["halo"]) > 100
    _check_exact(res, cases.building(**kw))


def test_world_one_is_the_plain_path():
    from buildingsegment_b200 import slabs

    xyz = cases.building(n=20000, order="shuffled")
    r = slabs.segment_tile(OracleBackend(), torch.from_numpy(xyz))
    P = O.pipeline(xyz)
    assert r["n_halo"] == 0 and r["n_planes"] == P["grow"].n_planes
    assert np.array_equal(r["labels"].numpy(), P["grow"].label)
    assert np.array_equal(r["plane_idx"].numpy(), P["grow"].plane_idx)


@pytest.mark.parametrize("world,case,kw", [(2, "building", dict(n=30000, order="shuffled")), (3, "block", dict(n=40000))])
def test_tile_raster_is_the_undivided_raster(world, case, kw, tmp_path):
    """The stitched pixel columns of the slabs == the raster of the undivided tile, bit for bit: doubles (the libm log
    of the count channel included), ground threshold and the save_image bytes; the faces are not bin-aligned."""
    res = _run(case, kw, 500, tmp_path, world=world, want_raster=True)
    xyz = getattr(cases, case)(**kw)
    xs, mn, mx, wh = O.bbox_shift(xyz)
    W, H = int(wh[0]), int(wh[1])
    th = O.orc().orc_ground_th(xs, len(xs), int(mx[2] - mn[2]), 1000)
    col = 0
    for r in res:
        assert (int(r["W"]), int(r["H"])) == (W, H) and float(r["th"]) == th
        assert int(r["x0"]) == col
        col += r["image"].shape[1]
    ref = O.raster(xs, mx[2] - mn[2], W, H)
    oa, ob, _, _ = O.save_image(ref)
    img = np.concatenate([r["image"] for r in res], axis=1)
    assert img.shape == ref.shape
    assert np.array_equal(img.view(np.int64), ref.view(np.int64))
    assert np.array_equal(np.concatenate([r["a"] for r in res], axis=1), oa)
    assert np.array_equal(np.concatenate([r["b"] for r in res], axis=1), ob)
    assert all(r["image"].shape[1] > 3 for r in res) and np.count_nonzero(img[..., 1]) > 0
    assert any(int(c) % 100 != 0 for c in res[0]["cuts"] - int(xyz[:, 0].min()))
