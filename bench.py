#!/usr/bin/env python
"""bench.py -- points/s of the segmentation hot path (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C5] [--points P]

A step = one pass of the whole hot path over one synthetic cloud:
  bbox+shift -> Morton binning -> exact kNN + PCA normals -> order-faithful plane growing -> labels ->
  ground threshold + height/count raster (count channel finalised on the host with libm, overlapped with the grower).
Default workload: C5, the 200 M-point synthetic city tile BASELINE.json's target is quoted on (it fits one B200);
C1..C4 are selectable (--workload) with their own parameters (WORKLOADS below).

`value`  : cloud already in HBM when the timed region starts (bseg_set_points_device + bseg_run_device)
`e2e`    : through bseg_segment_host with pinned HOST buffers: H2D of the cloud, D2H of the shifted cloud, the labels
           and the two PNG byte images inside the timed region (the same outputs at every N)
N > 1    : one process per GPU (torchrun), STRONG scaling on the same tile: every rank holds one contiguous chunk of the
           tile's points; partitioner (all_to_all) -> halo exchange -> kNN/normals per slab -> rows gathered on rank 0,
           which grows the undivided tile (labels exact by construction) -> labels scattered -> the tile's raster per
           slab (buildingsegment_b200/slabs.py, DESIGN.md "multi-GPU")
--impl reference : the CPU path (oracle/_ref = the reference's own grower/raster lines, oracle port for the Open3D
           kNN/normals the reference links but does not vendor) on the host cores, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# algorithmic (compulsory) bytes per point, SURVEY.md 8(d) / DESIGN.md "roofline" (K = 15; K = 16 adds 4 to knn)
BYTES_PER_POINT = {
    "bbox_keys": 12 + 12 + 12,      # bbox read, shift read+write
    "sort": 12 + 12 + 16 + 4,       # keys write, gather: read xyz, write pts + inv (radix passes are internal)
    "cells": 8 + 4,
    "knn": 12 + 60 + 24 + 8,        # read pts, write K=15 row, normal, curvature
    "grow": 106,                    # S4: row 60 + pos 12 + normal 24 + label 4 + colour 6, per point
    "raster": 16,
}

# BASELINE.json configs -> generator + parameters; `n` = the config's full size
WORKLOADS = {
    "C1": dict(n=1_000_000, params={}, what="single gable-roof building, reference defaults"),
    "C2": dict(n=10_000_000, params=dict(grow_radius=250.0),
               what="suburban block (40 buildings + ground + clutter), radius-search growing (grow_radius 250 mm)"),
    "C3": dict(n=50_000_000, params=dict(K=16, radius=1000.0),
               what="aerial tile, flight-line order, 16-bit quantised coordinates, K=16, radius 1000 mm"),
    "C4": dict(n=100_000_000, params=dict(K=16, radius=2.5, max_nn=50, th_thickness=3, bin=4, bin_height=16),
               what="voxelised dense scans (10-bit lattices, voxel units), K=16, raster output on"),
    "C5": dict(n=200_000_000, params={},
               what="city tile 2 x 2 km = 16 suburban blocks, block-major shuffled, reference defaults "
                    "(K=15, r=100 mm, max_nn=50, 300 mm, 0.88, 400)"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# workloads
def labels_checksum_parts(labels, offset):
    """(sum of labels, sum of labels x weight(global index)) mod 2^64 of a contiguous chunk of the tile that starts at
    global index `offset`, as two int64 (two's complement of the uint64 sums): chunks add up -- rank by rank through an
    int64 all_reduce -- to the checksum of the undivided tile, whatever the split."""
    s1 = np.uint64(0)
    s2 = np.uint64(0)
    n = len(labels)
    with np.errstate(over="ignore"):
        for a in range(0, n, 1 << 24):
            lab = labels[a: a + (1 << 24)].astype(np.uint64)
            w = np.arange(offset + a + 1, offset + a + 1 + len(lab), dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
            s1 = s1 + lab.sum(dtype=np.uint64)
            s2 = s2 + (lab * w).sum(dtype=np.uint64)

    def signed(v):
        v = int(v)
        return v - (1 << 64) if v >= (1 << 63) else v

    return signed(s1), signed(s2)


def make_chunk(name, n, rank=0, world=1, workers=1):
    """Rank `rank`'s contiguous chunk of workload `name` at `n` points (int32, unshifted); world == 1: the whole
    cloud.  Call before CUDA is initialised in this process (the block generators fork workers)."""
    from buildingsegment_b200 import synth

    if name == "C5":
        nb = 16
        if nb % world == 0:
            per = nb // world
            return synth.city_tile(n, blocks=range(rank * per, (rank + 1) * per), workers=workers)
        xyz = synth.city_tile(n, workers=workers)
    elif name == "C4":
        xyz = synth.voxel_scan(n, workers=workers)
    else:
        xyz = synth.make(name, n)
    if world == 1:
        return xyz
    cuts = [(len(xyz) * r) // world for r in range(world + 1)]
    return np.ascontiguousarray(xyz[cuts[rank]: cuts[rank + 1]])


def make_sample(name, n_full, target):
    """A bounded sample of the workload at the workload's own density, for the CPU legs: C5 -> its first blocks,
    cropped; the others -> a spatial crop of a (smaller) cloud of the same generator."""
    from buildingsegment_b200 import synth

    if name == "C5":
        counts = synth.city_tile_counts(n_full)
        nb = max(1, min(16, int(np.ceil(target / counts[0]))))
        xyz = synth.city_tile(n_full, blocks=range(nb), workers=nb)
        what = f"blocks 0..{nb - 1} of the tile ({len(xyz)} points)"
    elif name == "C4":
        xyz = synth.voxel_scan(min(n_full, max(target, 1_000_000)), workers=4)
        what = f"the first {len(xyz)} voxels of the scan"
    else:
        xyz = synth.make(name, min(n_full, 12_000_000))
        what = f"{name} generated at {len(xyz)} points"
    if len(xyz) > 1.2 * target:
        xyz = crop_sample(xyz, target)
        what += f", corner crop of {len(xyz)} points (same density)"
    return xyz, what


def crop_sample(xyz, target):
    """Spatial crop (same density as the full cloud): the corner square holding ~target points."""
    if len(xyz) <= target:
        return xyz
    x = xyz[:, 0] - xyz[:, 0].min()
    y = xyz[:, 1] - xyz[:, 1].min()
    frac = np.sqrt(target / len(xyz))
    side = frac * max(x.max(), y.max())
    for _ in range(20):
        m = (x < side) & (y < side)
        k = int(m.sum())
        if k < 0.8 * target:
            side *= 1.1
        elif k > 1.2 * target:
            side *= 0.95
        else:
            break
    return np.ascontiguousarray(xyz[m])


def cpu_path(xyz, use_ref, kw):
    """The CPU path on one cloud: oracle port for kNN/normals (OpenMP, all cores), the reference's own
    grower/raster lines when oracle/_ref is built (use_ref) else the oracle port.  Returns seconds."""
    import oracle_lib as O

    p = dict(O.DEFAULTS)
    p.update(kw)
    t0 = time.perf_counter()
    xs, mn, mx, wh = O.bbox_shift(xyz, p["bin"])
    kq = max(p["max_nn"], p["K"])
    idx, d2 = O.knn(xs, kq, cell=max(1, int(p["radius"])))
    nrm, _, _ = O.normals(xs, idx, d2, p["radius"], p["max_nn"])
    neigh = np.ascontiguousarray(idx[:, : p["K"]])
    if p["grow_radius"] > 0:
        cut = d2[:, : p["K"]].astype(np.float64) >= p["grow_radius"] ** 2
        cut[:, 0] = False
        neigh = np.ascontiguousarray(np.where(cut, -1, neigh).astype(np.int32))
    plain = p["K"] == 15 and p["th_thickness"] == 300 and p["bin"] == 100  # the literals the reference lines hard-code
    if use_ref and plain:
        O.ref_grow(xs, nrm, neigh)
        O.ref_raster(xyz)
    else:
        O.grow(xs, nrm, neigh, p["K"], p["th_thickness"], p["th_point_count"], p["th_dot"])
        img = O.raster(xs, mx[2] - mn[2], int(wh[0]), int(wh[1]), bin=p["bin"], bin_height=p["bin_height"], bias=p["count_bias"])
        O.save_image(img)
    return time.perf_counter() - t0, (use_ref and plain)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O

    O.orc()
    have_ref = O.ref() is not None
    cores = os.cpu_count() or 1
    W = WORKLOADS[args.workload]
    n_full = args.points or W["n"]
    xyz, what = make_sample(args.workload, n_full, args.ref_sample)
    times, used_ref = [], False
    for it in range(args.warmup + args.steps):
        dt, used_ref = cpu_path(xyz, have_ref, W["params"])
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = len(xyz) * len(times) / total
    kind = "reference" if used_ref else "port"
    sample = (f"{what}; kNN/normals = oracle port (Open3D 0.19 is not vendored), OpenMP x{cores}; grower+raster = "
              + ("the reference's own lines (oracle/_ref: O(P^2) per plane, hence the small sample), single thread as in the reference"
                 if used_ref else "oracle port, single thread"))
    out = {
        "impl": "reference", "metric": "points/sec segmented end-to-end", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {W['what']}, {n_full} points (CPU arm timed on a bounded sample: {len(xyz)} points)"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------------
def io_block(xyz, W, H):
    """ply::read / ply::write / PNG encode, timed by the host tool on a bounded sample (reported separately)."""
    exe = os.path.join(ROOT, "io_bench_b200")
    if not os.path.exists(exe):
        return None
    n = min(len(xyz), 5_000_000)
    rec = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    rec["x"], rec["y"], rec["z"] = (xyz[:n, k].astype(np.float32) / 1000.0 for k in range(3))
    try:
        with tempfile.TemporaryDirectory() as d:
            src, dst = os.path.join(d, "in.ply"), os.path.join(d, "out.ply")
            with open(src, "wb") as f:
                f.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {n}\nproperty float x\nproperty float y\n"
                         "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").encode())
                rec.tofile(f)
            w, h = min(int(W), 4000), min(int(H), 4000)
            r = subprocess.run([exe, src, dst, str(w), str(h)], capture_output=True, text=True, timeout=300)
            if r.returncode != 0:
                return {"error": r.stderr[-200:]}
            j = json.loads(r.stdout.strip().splitlines()[-1])
        return {"sample_points": n, "ply_read_points_per_s": n / j["ply_read_s"], "ply_write_points_per_s": n / j["ply_write_s"],
                "png_encode_mb_per_s": 3e-6 * w * h / j["png_encode_s"], "png_image": [w, h], "png_bytes": j["png_bytes"],
                "note": "host I/O either side of the path, one thread per file; not inside value / e2e"}
    except Exception as e:  # the io block is informational: never fail the bench line over it
        return {"error": str(e)[:200]}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    Wl = WORKLOADS[args.workload]
    n_full = args.points or Wl["n"]
    cores = os.cpu_count() or 1
    # synthetic input first: the block generators fork, which must happen before CUDA comes up in this process
    t_gen = time.perf_counter()
    xyz = make_chunk(args.workload, n_full, rank, world, workers=max(1, min(16, cores // max(1, world))))
    t_gen = time.perf_counter() - t_gen
    cpu_sample = make_sample(args.workload, n_full, args.cpu_sample) if (rank == 0 and not args.no_cpu) else None

    import torch
    import torch.distributed as dist

    from buildingsegment_b200 import lib

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up (NCCL_DEBUG=VERSION on the
        # boxes): stdout carries the ONE JSON line of the contract, so the banner goes to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            torch.cuda.set_device(local)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    n = len(xyz)
    ctx = lib.Context(local)
    p = lib.default_params(**Wl["params"])
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    d_xyz = torch.from_numpy(xyz).to(dev)
    n_tile = n
    if world > 1:
        from buildingsegment_b200 import slabs

        backend = slabs.CudaBackend(ctx, p)
        cnt = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        n_tile = int(cnt.item())
    info = {}
    phase = {}

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def tile_step(d_chunk):
        halo0 = args.halo or int(max(20 * p.radius, 2 * p.bin))
        # the raster needs the slabs, not the labels: every rank rasters its columns before the root starts growing
        seg = slabs.segment_tile(backend, d_chunk, halo=halo0, radius=p.radius,
                                 between=lambda prep: slabs.raster_tile(backend, d_chunk, prep, p.bin, p.bin_height, p.count_bias,
                                                                        want_image=False))
        img = seg["between"]
        info.update(n_planes=seg["n_planes"], halo=seg["halo"], n_halo=seg["n_halo"], raster_tile=[img["W"], img["H"]],
                    raster_columns=[img["x0"], img["x0"] + img["cols"]],
                    slab_points=int(seg["partition"].owned_counts[rank]))
        for k, v in seg["t"].items():
            phase[k] = phase.get(k, 0.0) + v
        return seg, img

    def step_device():
        if world == 1:
            ctx.set_points_device(d_xyz.data_ptr(), n)
            ctx.run_device(p, lib.RUN_ALL)
        else:
            tile_step(d_xyz)

    # ---- device-resident leg ----
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ctx.reset_counters()
    if world > 1 and backend.tile_ctx is not None:
        backend.tile_ctx.reset_counters()
    phase.clear()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    last_t = None
    e0.record(stream)
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        step_device()
        t = ctx.timings()
        if world > 1 and backend.tile_ctx is not None:  # the root's grower ran in the tile's context
            tt = backend.tile_ctx.timings()
            for k in tt:
                if k.startswith("grow") or k == "finalize":
                    t[k] = tt[k]
            t["kernel_launches"] += tt["kernel_launches"]
        for k in ("bbox_keys", "sort", "cells", "knn", "knn_fallback", "grow", "finalize", "raster", "raster_host"):
            stage_ms[k] = stage_ms.get(k, 0.0) + t[k]
        last_t = t
    e1.record(stream)
    barrier()
    ms_wall = (time.perf_counter() - t_wall) * 1e3
    clocks = sampler.stop()
    phase_dev = dict(phase)  # (the end-to-end leg below runs the same phases again)
    # one context's stream sees only its own kernels: at N > 1 the step also runs torch / NCCL work on other streams,
    # so the step time is the wall clock between the two barriers (device-synchronised on both sides)
    ms_dev = e0.elapsed_time(e1) if world == 1 else ms_wall
    launches = last_t["kernel_launches"]
    t_max = torch.tensor([ms_dev], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_dev = float(t_max.item())
    value = n_tile * args.steps / (ms_dev * 1e-3)

    # ---- end-to-end leg: pinned host buffers, the same outputs at every N ----
    if world == 1:
        W, H = ctx.raster_size(p)
    else:
        W, H = info["raster_tile"]
    h_xyz = torch.from_numpy(xyz).pin_memory()
    h_shift = torch.empty((n, 3), dtype=torch.int32).pin_memory()
    h_label = torch.empty(n, dtype=torch.int32).pin_memory()
    cols = W if world == 1 else info["raster_columns"][1] - info["raster_columns"][0]
    h_a = torch.empty((H, cols, 3), dtype=torch.uint8).pin_memory()
    h_b = torch.empty((H, cols, 3), dtype=torch.uint8).pin_memory()
    d_stage = torch.empty((n, 3), dtype=torch.int32, device=dev) if world > 1 else None
    npl_box = [0]

    def step_host():
        if world == 1:
            npl_box[0], _, _ = ctx.segment_host(p, h_xyz.numpy(), h_shift.numpy(), h_label.numpy(), h_a.numpy(), h_b.numpy())
            return
        d_stage.copy_(h_xyz, non_blocking=True)                       # H2D of this rank's chunk from pinned memory
        seg, img = tile_step(d_stage)
        origin = torch.from_numpy(seg["partition"].origin.astype(np.int32)).to(dev)
        h_shift.copy_(d_stage - origin[None, :], non_blocking=True)   # D2H: the chunk shifted to the tile origin (TMC3.cpp:71)
        h_label.copy_(seg["labels"], non_blocking=True)               # D2H: labels of the chunk
        h_a.copy_(img["png_a"], non_blocking=True)                    # D2H: this rank's columns of images A and B
        h_b.copy_(img["png_b"], non_blocking=True)
        torch.cuda.synchronize(dev)
        npl_box[0] = seg["n_planes"]

    for _ in range(max(1, min(2, args.warmup // 2))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize(dev)
    ms_e2e = (time.perf_counter() - t0) * 1e3
    t_max = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_max.item())
    e2e_value = n_tile * args.steps / (ms_e2e * 1e-3)
    # ---- checksum of the labels of the last end-to-end pass (outside the timed region): sum of labels and a
    #      position-weighted sum mod 2^64 over the TILE's index order -- the same number for every build, engine and N ----
    off = 0
    if world > 1:
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev))
        off = int(sum(int(c.item()) for c in counts[:rank]))
    s1, s2 = labels_checksum_parts(h_label.numpy(), off)
    chk = torch.tensor([s1, s2], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(chk)
    labels_checksum = [int(chk[0].item()) & ((1 << 64) - 1), "%016x" % (int(chk[1].item()) & ((1 << 64) - 1))]
    h2d = n_tile * 12
    d2h = n_tile * 12 + n_tile * 4 + 2 * 3 * W * H

    if rank == 0:
        peak, peak_src = peaks()
        per_step = {k: v / args.steps for k, v in stage_ms.items()}
        npts = {k: n_tile for k in per_step}
        if world > 1:  # sharded stages ran on this rank's slab (+ halo) only
            for k in ("bbox_keys", "sort", "cells", "knn", "knn_fallback", "raster"):
                npts[k] = info.get("slab_points", n) + info.get("n_halo", 0)
        stages = {}
        for k, ms in per_step.items():
            b = BYTES_PER_POINT.get(k)
            if b and ms > 0:
                gbs = b * npts[k] / (ms * 1e-3) / 1e9
                stages[k] = {"ms": round(ms, 3), "bytes_per_point": b, "points": npts[k], "gbs": round(gbs, 2), "frac": round(gbs / peak, 5)}
            else:
                stages[k] = {"ms": round(ms, 3)}
        dev_stages = {k: v for k, v in per_step.items() if k != "raster_host"}
        dom = max(dev_stages, key=lambda k: dev_stages[k])
        # Dominant stage: the plane grower.  Algorithmic bytes = SURVEY 8(d) S4, 106 B per point of the cloud, once --
        # released speculative work is NOT counted as useful (it is reported as speculation_overhead).
        grow_ms = per_step.get("grow", 0.0)
        ach = BYTES_PER_POINT["grow"] * n_tile / (grow_ms * 1e-3) / 1e9 if grow_ms > 0 else None
        sweep_ms, slice_ms = float(last_t["grow_sweep_ms"]), float(last_t["grow_slice_ms"])
        rounds = max(1, int(last_t["grow_rounds"]))
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tj = json.load(f)
                if tj.get("workload_key") == args.workload and n_tile == 10_000_000:  # measured on that workload only
                    traffic = tj.get("grow_dram_bytes_per_step") / rounds   # per launch (= round), like avg_launch_ms
        except Exception:
            pass
        committed = int(last_t["grow_steps"]) - int(last_t["grow_tiny_tx"])
        roofline = {"bound": "hbm", "kernel": "plane grower: spec_sweep_kernel (sequential authority) + spec_grow_kernel (speculative "
                                              "slices); dependent-gather latency bound, see DESIGN.md 6",
                    "achieved": round(ach, 3) if ach else None, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": round(ach / peak, 6) if ach else None, "traffic": traffic,
                    "algorithmic_bytes": BYTES_PER_POINT["grow"] * n_tile, "bytes_per_point": BYTES_PER_POINT["grow"],
                    "launches": rounds, "avg_launch_ms": round(grow_ms / rounds, 4),
                    "kernels_ms": {"spec_sweep_kernel+apply": round(sweep_ms, 2), "spec_grow_kernel": round(slice_ms, 2)},
                    "speculation_overhead": {"released_calls": int(last_t["grow_wasted_steps"]), "committed_plane_calls": committed},
                    "dominant_stage": dom, "stages": stages}
        cpu = None
        if not args.no_cpu:
            import oracle_lib as O

            O.orc()
            sample, what = cpu_sample
            dt, _ = cpu_path(sample, False, Wl["params"])
            cpu = {"value": len(sample) / dt, "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": f"{what}, one pass, {dt:.1f} s; kNN/normals OpenMP on all cores, grower/raster single thread "
                             f"(as the reference)"}
        io = None if args.no_io else io_block(xyz, W, H)
        config = {"workload": f"{args.workload}: {Wl['what']}; {n_tile} points" + (f" over {world} x-slabs" if world > 1 else ""),
                  "points": n_tile, "params": Wl["params"], "l2": "inputs larger than L2 (cloud + neighbour rows >> 126 MB)",
                  "grow_engine": "sweeper + speculative growers (grow_mode 0)", "planes": int(npl_box[0]),
                  "labels_checksum": labels_checksum,
                  "raster": [int(W), int(H)], "generate_s": round(t_gen, 1)}
        if world > 1:
            config["tile"] = {**{k: info[k] for k in ("halo", "n_halo", "slab_points", "raster_columns")},
                              "phases_ms_per_step_rank0": {k: round(1e3 * v / args.steps, 2) for k, v in phase_dev.items()},
                              "labels": "exact: rank 0 grows the undivided tile from the slabs' rows / normals"}
        out = {
            "metric": "points/sec segmented end-to-end", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": config,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "io": io,
            "grow": {"steps": int(last_t["grow_steps"]), "rounds": int(last_t["grow_rounds"]),
                     "n_unresolved_knn": int(last_t["n_unresolved"]), "n_big_items": int(last_t["n_big_cells"]),
                     "wasted_steps": int(last_t["grow_wasted_steps"]), "sweep_iters": int(last_t["grow_sweep_iters"]),
                     "tiny_tx": int(last_t["grow_tiny_tx"]), "seq_fallbacks": int(last_t["grow_seq_fallbacks"]),
                     "head_steps": int(last_t["grow_head_steps"]), "head_ms": round(last_t["grow_head_ns"] / 1e6, 3),
                     "sweep_ms": round(last_t["grow_sweep_ns"] / 1e6, 3), "at_fails": int(last_t["grow_at_fails"])},
        }
        print(json.dumps(out))
    if world > 1:
        backend.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="0 = the config's full size")
    ap.add_argument("--cpu-sample", type=int, default=25_000_000)
    ap.add_argument("--ref-sample", type=int, default=60_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-io", action="store_true")
    ap.add_argument("--halo", type=int, default=0, help="initial slab halo width in cloud units (N > 1), 0 = 20 x the search "
                                                        "radius; doubled until sufficient")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
