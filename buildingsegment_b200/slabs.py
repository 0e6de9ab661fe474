"""Multi-GPU slabs: one process per GPU, the tile split along x (BASELINE.json config C5).

The reference (tmc3) is a single-threaded program; nothing in it shards.  What is sharded here, and what the
result means:

  1. tile origin      all_reduce(MIN) of the slab minima -> every rank shifts by the SAME origin
                      (TMC3.cpp:70-72 subtracts the minimum of the whole cloud; bseg_set_origin)
  2. halo exchange    every rank sends the points within `halo` of its two faces to the neighbour ranks
                      (torch.distributed P2P: NCCL over NVLink on GPUs, gloo in the CPU tests); the local cloud
                      is [owned points in their own order | halo from the left | halo from the right]
  3. kNN + normals    per rank on the local cloud.  For an owned point they equal the rows / normals of the
                      undivided tile iff its K-th neighbour is not farther than `halo` (and halo >= radius):
                      a missing point is at least `halo` away.  bseg_halo_check counts the violations; the halo
                      is doubled and the exchange repeated until there are none.
  4. plane growing    per rank on the local cloud (owned seeds first, halo seeds last).  The reference's grower
                      is a globally index-ordered greedy algorithm, so labels of the undivided tile are NOT
                      reproduced across a face; the per-slab result is exact for the slab's own cloud.
  5. label merge      planes of two ranks that contain the same physical point (an owned point and its halo
                      copy, both labelled) are the same surface: the pairs are all-gathered, a union-find gives
                      every plane the smallest global id of its component ("canonical minimum-index label").
                      Global id of local plane k on rank r = 1 + sum(planes of ranks < r) + (k - 1).

  6. raster           (raster_slab) the tile's height / count image, bit for bit: tile extent by all_reduce(MAX),
                      the tile's ground threshold from the summed z histograms (TMC3.cpp:181-198), every rank
                      rasters [left halo | owned | right halo] -- the tile's point order restricted to what can
                      reach its pixel columns, so the in-order fp64 sums of TMC3.cpp:127-172 are the tile's own --
                      keeps the pixel columns that start in its slab, and the per-channel maxima of save_image
                      (TMC3.cpp:81-121) come from an all_reduce(MAX).

The compute of steps 3-4 and 6 is a `backend` object: `CudaBackend` (libbseg, the product) or, in the CPU tests
only, a stand-in built on the oracle.  Nothing here falls back to a CPU path by itself.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------------------
# backends
class CudaBackend:
    """Steps 3-4 on this rank's GPU through the C ABI; tensors stay on the device."""

    def __init__(self, ctx, params):
        self.ctx = ctx
        self.p = params
        self.device = None

    def segment(self, xyz_local: torch.Tensor, n_owned: int, origin, x_lo: int, x_hi: int, halo: int):
        """xyz_local: int32 [n][3] CUDA tensor (unshifted).  Returns (label tensor [n] on the device,
        number of local planes, number of owned points whose neighbourhood may reach past the halo)."""
        from . import lib

        n = int(xyz_local.shape[0])
        ctx = self.ctx
        torch.cuda.current_stream(xyz_local.device).synchronize()
        ctx.set_origin(origin)
        ctx.set_points_device(xyz_local.data_ptr(), n)
        ctx.set_owned(n_owned)
        ctx.run_device(self.p, lib.RUN_KNN)
        bad = ctx.halo_check(x_lo - int(origin[0]), x_hi - int(origin[0]), halo)
        if bad:
            return None, 0, bad
        ctx.run_device(self.p, lib.RUN_GROW)
        d_label, _, _ = ctx.device_results()
        npl = ctx.n_planes()
        label = _wrap_device_int32(d_label, n, xyz_local.device)
        return label, npl, 0


    def raster(self, xyz_local: torch.Tensor, origin, ground_th: float):
        """Height / count image (doubles, before save_image) of `xyz_local` in ITS order with the tile's ground
        threshold: float64 [H][W][3] tensor on the device."""
        n = int(xyz_local.shape[0])
        ctx = self.ctx
        torch.cuda.current_stream(xyz_local.device).synchronize()
        ctx.set_origin(origin)
        ctx.set_points_device(xyz_local.data_ptr(), n)
        d_img, W, H = ctx.raster_device(self.p, ground_th)

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (H, W, 3), "typestr": "<f8", "data": (int(d_img), False), "version": 2}
        return torch.as_tensor(h, device=xyz_local.device).clone()


def _wrap_device_int32(ptr: int, n: int, device):
    """A torch view of library-owned device memory (valid until the next set_points)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=device)


# ---------------------------------------------------------------------------------------------------------
# step 2: halo exchange
def exchange_halo(xyz_owned: torch.Tensor, x_lo: int, x_hi: int, halo: int, group=None):
    """Returns (halo points from the left neighbour, from the right neighbour, their indices in the owner's
    cloud): int32 tensors on xyz_owned's device.  Rank r owns x in [x_lo, x_hi)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = xyz_owned.device
    empty = torch.empty((0, 4), dtype=torch.int32, device=dev)
    if world == 1:
        return empty, empty
    x = xyz_owned[:, 0]
    idx = torch.arange(xyz_owned.shape[0], dtype=torch.int32, device=dev)
    send_l = send_r = empty
    if rank > 0:
        m = x < x_lo + halo
        send_l = torch.cat([xyz_owned[m], idx[m, None]], dim=1).contiguous()
    if rank < world - 1:
        m = x >= x_hi - halo
        send_r = torch.cat([xyz_owned[m], idx[m, None]], dim=1).contiguous()
    # sizes first (two ints per rank), then the payloads as P2P operations
    counts = torch.tensor([send_l.shape[0], send_r.shape[0]], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    recv_l = recv_r = empty
    ops = []
    if rank > 0:
        recv_l = torch.empty((int(all_counts[rank - 1][1]), 4), dtype=torch.int32, device=dev)
        if send_l.numel():
            ops.append(dist.P2POp(dist.isend, send_l, rank - 1, group))
        if recv_l.numel():
            ops.append(dist.P2POp(dist.irecv, recv_l, rank - 1, group))
    if rank < world - 1:
        recv_r = torch.empty((int(all_counts[rank + 1][0]), 4), dtype=torch.int32, device=dev)
        if send_r.numel():
            ops.append(dist.P2POp(dist.isend, send_r, rank + 1, group))
        if recv_r.numel():
            ops.append(dist.P2POp(dist.irecv, recv_r, rank + 1, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv_l, recv_r


# ---------------------------------------------------------------------------------------------------------
# step 5: label merge
def _union_find_min(n_ids: int, pairs: np.ndarray) -> np.ndarray:
    """canonical[id] = smallest id of id's component (ids 1..n_ids; index 0 = unlabelled stays 0)."""
    parent = np.arange(n_ids + 1, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in pairs:
        ra, rb = find(int(a)), find(int(b))
        if ra != rb:
            if ra < rb:
                parent[rb] = ra
            else:
                parent[ra] = rb
    return np.array([find(i) for i in range(n_ids + 1)], dtype=np.int64)


def merge_labels(label_local: torch.Tensor, n_owned: int, n_planes: int, halo_l: torch.Tensor, halo_r: torch.Tensor,
                 group=None):
    """label_local: labels (0 or 1..n_planes, local ids) of [owned | halo_l | halo_r].  Returns
    (global canonical labels of the owned points, total planes over all ranks, components after the merge)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = label_local.device
    if world == 1:
        return label_local[:n_owned].to(torch.int64), n_planes, n_planes
    npl = torch.tensor([n_planes], dtype=torch.int64, device=dev)
    all_npl = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_npl, npl, group=group)
    counts = [int(t.item()) for t in all_npl]
    offs = np.concatenate([[0], np.cumsum(counts)])
    total = int(offs[-1])
    my_off = int(offs[rank])
    lab = label_local.to(torch.int64)
    glob = torch.where(lab > 0, lab + my_off, lab)  # global ids, 0 stays 0
    nl, nr = halo_l.shape[0], halo_r.shape[0]
    # tell the owner what I call its points: (index in the owner's cloud, my global id); labelled copies only
    def pack(h, sl):
        g = glob[sl]
        m = g > 0
        return torch.stack([h[:, 3].to(torch.int64)[m], g[m]], dim=1).contiguous()

    to_l = pack(halo_l, slice(n_owned, n_owned + nl)) if nl else torch.empty((0, 2), dtype=torch.int64, device=dev)
    to_r = pack(halo_r, slice(n_owned + nl, n_owned + nl + nr)) if nr else torch.empty((0, 2), dtype=torch.int64, device=dev)
    cnt = torch.tensor([to_l.shape[0], to_r.shape[0]], dtype=torch.int64, device=dev)
    all_cnt = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_cnt, cnt, group=group)
    ops, from_l, from_r = [], None, None
    if rank > 0:
        from_l = torch.empty((int(all_cnt[rank - 1][1]), 2), dtype=torch.int64, device=dev)
        if to_l.numel():
            ops.append(dist.P2POp(dist.isend, to_l, rank - 1, group))
        if from_l.numel():
            ops.append(dist.P2POp(dist.irecv, from_l, rank - 1, group))
    if rank < world - 1:
        from_r = torch.empty((int(all_cnt[rank + 1][0]), 2), dtype=torch.int64, device=dev)
        if to_r.numel():
            ops.append(dist.P2POp(dist.isend, to_r, rank + 1, group))
        if from_r.numel():
            ops.append(dist.P2POp(dist.irecv, from_r, rank + 1, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    # an owned point labelled on both sides ties the two planes together
    pairs = []
    for f in (from_l, from_r):
        if f is None or f.numel() == 0:
            continue
        mine = glob[f[:, 0]]
        m = mine > 0
        pairs.append(torch.stack([mine[m], f[:, 1][m]], dim=1))
    pairs = torch.unique(torch.cat(pairs, dim=0), dim=0) if pairs else torch.empty((0, 2), dtype=torch.int64, device=dev)
    # all ranks learn all pairs (padded all_gather), then run the same union-find
    npairs = torch.tensor([pairs.shape[0]], dtype=torch.int64, device=dev)
    all_np = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_np, npairs, group=group)
    mx = max(int(t.item()) for t in all_np)
    all_pairs = np.empty((0, 2), np.int64)
    if mx > 0:
        pad = torch.zeros((mx, 2), dtype=torch.int64, device=dev)
        pad[: pairs.shape[0]] = pairs
        gathered = [torch.zeros((mx, 2), dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, pad, group=group)
        all_pairs = np.concatenate([g[: int(c.item())].cpu().numpy() for g, c in zip(gathered, all_np)], axis=0)
    canon = _union_find_min(total, all_pairs)
    n_comp = int(len(np.unique(canon[1:]))) if total else 0
    canon_t = torch.from_numpy(canon).to(dev)
    return canon_t[glob[:n_owned]], total, n_comp


# ---------------------------------------------------------------------------------------------------------
def tile_origin(xyz_owned: torch.Tensor, group=None):
    """Step 1: the minimum of the whole tile, identical on every rank."""
    mn = xyz_owned.min(dim=0).values.to(torch.int64) if xyz_owned.shape[0] else torch.full((3,), 2**31 - 1, dtype=torch.int64,
                                                                                        device=xyz_owned.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    return mn.to(torch.int32).cpu().numpy()


def segment_slab(backend, xyz_owned: torch.Tensor, x_lo: int, x_hi: int, halo: int = 500, group=None, max_tries: int = 4):
    """The whole multi-GPU pass for this rank's slab (x in [x_lo, x_hi), unshifted integer units).
    Returns a dict: labels (int64 [n_owned], canonical global plane ids, 0 = none), n_planes_total,
    n_components, halo (the width that was sufficient), n_halo (points received)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    origin = tile_origin(xyz_owned, group)
    n_owned = int(xyz_owned.shape[0])
    for _ in range(max_tries):
        hl, hr = exchange_halo(xyz_owned, x_lo, x_hi, halo, group)
        local = torch.cat([xyz_owned, hl[:, :3], hr[:, :3]], dim=0).contiguous()
        label, npl, bad = backend.segment(local, n_owned, origin, x_lo, x_hi, halo)
        flag = torch.tensor([bad], dtype=torch.int64, device=xyz_owned.device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) == 0:
            break
        halo *= 2  # some neighbourhood may reach past the halo on some rank: widen it everywhere
    else:
        raise RuntimeError(f"halo of {halo} units still insufficient after {max_tries} doublings")
    labels, total, ncomp = merge_labels(label, n_owned, npl, hl, hr, group)
    return {"labels": labels, "n_planes_total": total, "n_components": ncomp, "halo": halo,
            "n_halo": int(hl.shape[0] + hr.shape[0]), "n_planes_local": npl, "halo_l": hl, "halo_r": hr, "origin": origin}


# ---------------------------------------------------------------------------------------------------------
# step 6: the tile's raster
def ground_threshold(xyz_owned: torch.Tensor, origin, zext: int, bin_height: int, group=None) -> int:
    """buildingSeg::groundTH (TMC3.cpp:181-198) of the whole tile: z histogram in bins of bin_height summed over
    the ranks, first bin whose cumulative count exceeds N/2 (N/2 in integers), times bin_height."""
    dev = xyz_owned.device
    nb = int(zext) // int(bin_height) + 1
    z = (xyz_owned[:, 2].to(torch.int64) - int(origin[2])) // int(bin_height)
    hist = torch.bincount(z, minlength=nb).to(torch.int64)[:nb] if xyz_owned.shape[0] else torch.zeros(nb, dtype=torch.int64, device=dev)
    total = torch.tensor([xyz_owned.shape[0]], dtype=torch.int64, device=dev)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    th_count = int(total.item()) // 2
    cum = torch.cumsum(hist, 0)
    over = torch.nonzero(cum > th_count)
    i = int(over[0].item()) if over.numel() else nb
    return i * int(bin_height)


def raster_slab(backend, xyz_owned: torch.Tensor, halo_l: torch.Tensor, halo_r: torch.Tensor, origin, x_lo: int, x_hi: int,
                halo: int, bin: int = 100, bin_height: int = 1000, group=None):
    """This rank's pixel columns of the tile's raster.  Returns a dict: image (float64 [H][cols][3], the doubles of
    compute_gird_picture), png_a / png_b (uint8 [H][cols][3], the bytes of save_image), x0 (first column), W, H
    (the tile's), ground_th, maxima (the tile's per-channel maxima)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = xyz_owned.device
    if world > 1 and halo < 2 * bin:
        raise ValueError("raster_slab: the halo must cover two raster bins")
    # the tile's extent: width / height as TMC3.cpp:75-76, z extent for the ground threshold
    mx = xyz_owned.max(dim=0).values.to(torch.int64) if xyz_owned.shape[0] else torch.full((3,), -2**31, dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    ext = [int(mx[k].item()) - int(origin[k]) for k in range(3)]
    W, H = ext[0] // bin + 2, ext[1] // bin + 2
    th = ground_threshold(xyz_owned, origin, ext[2], bin_height, group)
    # the tile's point order is rank-major: left halo (the left neighbour's points, in its order), owned, right halo
    local = torch.cat([halo_l[:, :3], xyz_owned, halo_r[:, :3]], dim=0).contiguous()
    img = backend.raster(local, origin, float(th))  # [H_l][W_l][3]
    # pixel columns that START in this slab (a column on a face goes to the rank on its left edge)
    c0 = 0 if rank == 0 else -((-(x_lo - int(origin[0]))) // bin)
    c1 = W if rank == world - 1 else -((-(x_hi - int(origin[0]))) // bin)
    c1 = max(c0, min(c1, W))
    block = torch.zeros((H, c1 - c0, 3), dtype=torch.float64, device=dev)
    hl_, wl_ = int(img.shape[0]), int(img.shape[1])
    hh, ww = min(H, hl_), min(c1, wl_)
    if ww > c0 and hh > 0:
        block[:hh, : ww - c0] = img[:hh, c0:ww]
    # save_image (TMC3.cpp:81-121): per-channel maximum over the tile, byte = (uint8)(255.0 * (v / max))
    m = block.reshape(-1, 3).max(dim=0).values if block.numel() else torch.zeros(3, dtype=torch.float64, device=dev)
    m = torch.clamp(m, min=0.0)
    if world > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    png_a = torch.zeros(block.shape, dtype=torch.uint8, device=dev)
    png_b = torch.zeros(block.shape, dtype=torch.uint8, device=dev)
    if float(m[0]) != 0.0:
        png_a[..., 0] = (255.0 * (1.0 * block[..., 0] / m[0])).to(torch.uint8)
    if float(m[1]) != 0.0:
        png_b[..., 1] = (255.0 * (1.0 * block[..., 1] / m[1])).to(torch.uint8)
    return {"image": block, "png_a": png_a, "png_b": png_b, "x0": c0, "W": W, "H": H, "ground_th": th, "maxima": m}
