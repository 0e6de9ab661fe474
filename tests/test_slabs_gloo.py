"""Multi-GPU slab logic (buildingsegment_b200/slabs.py) on the CPU: world_size 2 over gloo.

The compute backend here is a stand-in built on the CPU oracle (test infrastructure); what is under test is the
host logic of the N > 1 path: tile origin, halo exchange, halo sufficiency, the cross-slab label merge.
  * kNN rows and normals of every OWNED point equal those of the undivided cloud (bit for bit);
  * a plane that crosses the slab face ends up with ONE canonical (minimum) global id on both ranks;
  * plane counts add up and the merge only ever lowers ids.
"""
import os
import socket

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
        self.K, self.radius, self.max_nn = K, radius, max_nn
        self.last = None

    def segment(self, xyz_local, n_owned, origin, x_lo, x_hi, halo):
        xs = np.ascontiguousarray(xyz_local.numpy().astype(np.int64) - np.asarray(origin, np.int64)).astype(np.int32)
        idx, d2 = O.knn(xs, self.max_nn, cell=100)
        nrm, _, _ = O.normals(xs, idx, d2, self.radius, self.max_nn)
        neigh = np.ascontiguousarray(idx[:, : self.K])
        # halo sufficiency, as bseg_halo_check: owned points near a face whose K-th neighbour is beyond the halo
        x = xs[:n_owned, 0].astype(np.int64)
        near = (x - (x_lo - int(origin[0])) < halo) | ((x_hi - int(origin[0])) - x <= halo)
        dk = d2[:n_owned, self.K - 1]
        bad = int(np.count_nonzero(near & ((dk < 0) | (dk > halo * halo))))
        if bad:
            return None, 0, bad
        g = O.grow(xs, nrm, neigh)
        self.last = dict(xs=xs, neigh=neigh, nrm=nrm, grow=g)
        return torch.from_numpy(g.label.astype(np.int32)), int(g.n_planes), 0

    def raster(self, xyz_local, origin, ground_th):
        xs = np.ascontiguousarray(xyz_local.numpy().astype(np.int64) - np.asarray(origin, np.int64)).astype(np.int32)
        W, H = int(xs[:, 0].max()) // 100 + 2, int(xs[:, 1].max()) // 100 + 2
        return torch.from_numpy(O.raster_th(xs, ground_th, W, H))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, kw, halo, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from buildingsegment_b200 import slabs

        xyz = getattr(cases, case)(**kw)
        gid = np.arange(len(xyz))
        qs = np.quantile(xyz[:, 0], [k / world for k in range(1, world)]).astype(np.int64)
        edges = [int(xyz[:, 0].min())] + [int(q) for q in qs] + [int(xyz[:, 0].max()) + 1]
        lo, hi = edges[rank], edges[rank + 1]
        m = (xyz[:, 0] >= lo) & (xyz[:, 0] < hi)
        owned = torch.from_numpy(np.ascontiguousarray(xyz[m]))
        be = OracleBackend()
        r = slabs.segment_slab(be, owned, lo, hi, halo=halo)
        # local index -> global index: owned first, then the halo copies (re-run the exchange to learn them)
        hl, hr = slabs.exchange_halo(owned, lo, hi, r["halo"])
        gid_own = gid[m]
        other = [None] * world
        dist.all_gather_object(other, gid_own)
        gl = other[rank - 1][hl.numpy()[:, 3].astype(np.int64)] if rank > 0 and len(hl) else np.empty(0, np.int64)
        gr = other[rank + 1][hr.numpy()[:, 3].astype(np.int64)] if rank < world - 1 and len(hr) else np.empty(0, np.int64)
        l2g = np.concatenate([gid_own, gl, gr])
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), gid_own=gid_own, l2g=l2g, neigh=be.last["neigh"][: len(gid_own)],
                 nrm=be.last["nrm"][: len(gid_own)], labels=r["labels"].numpy(), total=r["n_planes_total"],
                 ncomp=r["n_components"], n_local=r["n_planes_local"], halo=r["halo"], n_halo=r["n_halo"],
                 local_label=be.last["grow"].label[: len(gid_own)])
    finally:
        dist.destroy_process_group()


def _run(case, kw, halo, tmp_path, world=2):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, kw, halo, str(tmp_path)), nprocs=world, join=True)
    return [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]


@pytest.mark.parametrize("case,kw", [("building", dict(n=30000, order="shuffled"))])
def test_two_slabs_match_undivided_knn_and_merge(case, kw, tmp_path):
    res = _run(case, kw, 400, tmp_path)
    xyz = getattr(cases, case)(**kw)
    P = O.pipeline(xyz)  # the undivided cloud (shifted by its own minimum == the tile origin)
    n_lab = 0
    for r in res:
        g = r["gid_own"]
        # neighbour rows: local indices -> global ids must equal the undivided rows, ties included
        rows = r["l2g"][r["neigh"]]
        assert np.array_equal(rows, P["neigh"][g]), "kNN rows of owned points differ from the undivided cloud"
        assert np.array_equal(r["nrm"].view(np.int64), P["normals"][g].view(np.int64)), "normals differ"
        assert r["n_halo"] > 0
        n_lab += int(np.count_nonzero(r["labels"]))
    total = int(res[0]["total"])
    assert total == int(res[0]["n_local"]) + int(res[1]["n_local"]) == int(res[1]["total"])
    assert int(res[0]["ncomp"]) == int(res[1]["ncomp"]) <= total
    for r in res:  # canonical ids: never above the rank-offset local id, 0 stays 0
        off = 0 if r is res[0] else int(res[0]["n_local"])
        loc = r["local_label"].astype(np.int64)
        assert np.array_equal(r["labels"] == 0, loc == 0)
        assert np.all(r["labels"][loc > 0] <= loc[loc > 0] + off)
    assert n_lab > 0


def test_three_slabs_middle_rank_has_two_neighbours(tmp_path):
    """World 3: the middle rank exchanges halos on both faces; owned kNN rows / normals still equal the undivided
    cloud's (a cloud without exact distance ties: ties between an owned point and a halo copy go by local index)."""
    kw = dict(n=30000, order="shuffled")
    res = _run("building", kw, 400, tmp_path, world=3)
    P = O.pipeline(cases.building(**kw))
    for r in res:
        g = r["gid_own"]
        assert np.array_equal(r["l2g"][r["neigh"]], P["neigh"][g])
        assert np.array_equal(r["nrm"].view(np.int64), P["normals"][g].view(np.int64))
    assert int(res[1]["n_halo"]) > int(res[0]["n_halo"]) > 0  # two faces against one
    assert len({int(r["total"]) for r in res}) == 1 and len({int(r["ncomp"]) for r in res}) == 1


def test_plane_across_two_faces_gets_one_id(tmp_path):
    """One flat plane cut in three: the big plane carries the same canonical id on all three ranks."""
    res = _run("grid_plane", dict(nx=180, ny=50, order="shuffled"), 400, tmp_path, world=3)
    big = [np.bincount(r["labels"][r["labels"] > 0]).argmax() for r in res]
    assert big[0] == big[1] == big[2] == 1, big
    assert int(res[0]["ncomp"]) < int(res[0]["total"])


def test_insufficient_halo_is_doubled_until_it_suffices(tmp_path):
    """A halo narrower than some K-th neighbour distance: the sufficiency check (bseg_halo_check) fails on some rank,
    every rank doubles the halo and exchanges again; the rows of the owned points are the undivided cloud's."""
    kw = dict(n=30000, order="shuffled")
    res = _run("building", kw, 100, tmp_path)
    P = O.pipeline(cases.building(**kw))
    assert int(res[0]["halo"]) == int(res[1]["halo"]) > 100
    for r in res:
        g = r["gid_own"]
        assert np.array_equal(r["l2g"][r["neigh"]], P["neigh"][g])
        assert np.array_equal(r["nrm"].view(np.int64), P["normals"][g].view(np.int64))


def test_plane_across_the_face_gets_one_id(tmp_path):
    """One flat plane cut in two: every labelled point of the big plane carries the same canonical id on both ranks."""
    res = _run("grid_plane", dict(nx=120, ny=60, order="shuffled"), 400, tmp_path)
    ids = [np.unique(r["labels"][r["labels"] > 0]) for r in res]
    big = [np.bincount(r["labels"][r["labels"] > 0]).argmax() for r in res]
    assert big[0] == big[1] == 1, (big, ids)
    assert int(res[0]["ncomp"]) < int(res[0]["total"])


def test_world_one_is_the_plain_path():
    from buildingsegment_b200 import slabs

    xyz = cases.building(n=20000, order="shuffled")
    be = OracleBackend()
    r = slabs.segment_slab(be, torch.from_numpy(xyz), int(xyz[:, 0].min()), int(xyz[:, 0].max()) + 1)
    P = O.pipeline(xyz)
    assert r["n_halo"] == 0 and r["n_planes_total"] == P["grow"].n_planes
    assert np.array_equal(r["labels"].numpy(), P["grow"].label.astype(np.int64))


def test_union_find_canonical_minimum():
    from buildingsegment_b200.slabs import _union_find_min

    c = _union_find_min(6, np.array([[5, 2], [2, 6], [3, 4]]))
    assert c.tolist() == [0, 1, 2, 3, 3, 2, 2]


def _raster_worker(rank, world, port, case, kw, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from buildingsegment_b200 import slabs

        xyz = getattr(cases, case)(**kw)
        # faces that are not aligned with the raster bins; at world 3 the middle rank has two neighbours
        qs = np.quantile(xyz[:, 0], [k / world for k in range(1, world)]).astype(np.int64) + 37
        edges = [int(xyz[:, 0].min())] + [int(q) for q in qs] + [int(xyz[:, 0].max()) + 1]
        lo, hi = edges[rank], edges[rank + 1]
        m = (xyz[:, 0] >= lo) & (xyz[:, 0] < hi)
        owned = torch.from_numpy(np.ascontiguousarray(xyz[m]))
        origin = slabs.tile_origin(owned)
        hl, hr = slabs.exchange_halo(owned, lo, hi, 500)
        r = slabs.raster_slab(OracleBackend(), owned, hl, hr, origin, lo, hi, 500)
        np.savez(os.path.join(out_dir, f"raster{rank}.npz"), image=r["image"].numpy(), a=r["png_a"].numpy(), b=r["png_b"].numpy(),
                 x0=r["x0"], W=r["W"], H=r["H"], th=r["ground_th"], owned=owned.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case,kw", [(2, "building", dict(n=30000, order="shuffled")), (2, "block", dict(n=40000)),
                                           (3, "block", dict(n=40000))])
def test_slabs_raster_is_the_tiles_raster(world, case, kw, tmp_path):
    """The stitched pixel columns of the slabs == the raster of the undivided tile (points in rank-major order),
    bit for bit: doubles, ground threshold and the save_image bytes."""
    port = _free_port()
    mp.spawn(_raster_worker, args=(world, port, case, kw, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(os.path.join(tmp_path, f"raster{r}.npz")) for r in range(world)]
    tile = np.concatenate([r["owned"] for r in res], axis=0)  # the tile's point order is rank-major
    xs, mn, mx, wh = O.bbox_shift(tile)
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
