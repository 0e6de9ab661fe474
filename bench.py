#!/usr/bin/env python
"""bench.py -- points/s of the segmentation hot path (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points P] [--workload C2]

A step = one pass of the whole hot path over one synthetic cloud:
  bbox+shift -> Morton binning -> exact kNN + PCA normals -> order-faithful plane growing ->
  labels -> ground threshold + height/count raster.
`value`  : cloud already in HBM when the timed region starts (bseg_set_points_device + bseg_run_device)
`e2e`    : through bseg_segment_host with pinned HOST buffers, H2D of the cloud and D2H of the shifted
           cloud, labels and PNG bytes inside the timed region
N > 1    : one process per GPU (torchrun); every rank owns one 200 m C2-like x-slab of the city tile (weak scaling):
           shared tile origin, NCCL halo exchange with the neighbour ranks, halo sufficiency check, per-slab
           segmentation, cross-slab label merge (buildingsegment_b200/slabs.py, DESIGN.md "multi-GPU")
--impl reference : the CPU path (oracle/_ref = the reference's own grower/raster lines, oracle port for
           the Open3D kNN/normals the reference links but does not vendor) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# algorithmic (compulsory) bytes per point, SURVEY.md 8(d) / DESIGN.md "roofline"
BYTES_PER_POINT = {
    "bbox_keys": 12 + 12 + 12,      # bbox read, shift read+write
    "sort": 12 + 12 + 16 + 4,       # keys write, gather: read xyz, write pts + inv (radix passes are internal)
    "cells": 8 + 4,
    "knn": 12 + 60 + 24 + 8,        # read pts, write K=15 row, normal, curvature
    "grow": 106,                    # per visit: row 60 + pos 12 + normal 24 + state/label 10
    "raster": 16,
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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


SLAB_M = 200.0  # slab width of the multi-GPU tile, metres


def make_workload(name, n, rank=0, world=1):
    """Synthetic cloud of config `name` at `n` points per rank (int32 mm, unshifted)."""
    from buildingsegment_b200 import synth

    if world > 1:
        # C5-style city tile = C2-like blocks side by side: rank r owns the 200 m block at x = r * 200 m, so the
        # per-GPU work is the N = 1 workload (same generator, another seed) plus the exchange with the neighbours
        # (seed 1002 + rank: rank 0's block IS the C2 cloud of the N = 1 run, so the per-N values are comparable)
        rng = np.random.default_rng(1002 + rank)
        pts = synth._block(rng, n, rank * SLAB_M, 0.0, SLAB_M, 40, 0.15, "shuffled")
        return np.ascontiguousarray(synth.to_mm(pts))
    return synth.make(name, n)


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


def cpu_path(xyz, use_ref):
    """The CPU path on one cloud: oracle port for kNN/normals (OpenMP, all cores), the reference's own
    grower/raster lines when oracle/_ref is built (use_ref) else the oracle port.  Returns seconds."""
    import oracle_lib as O

    t0 = time.perf_counter()
    xs, mn, mx, wh = O.bbox_shift(xyz)
    idx, d2 = O.knn(xs, 50, cell=100)
    nrm, _, _ = O.normals(xs, idx, d2, 100.0, 50)
    neigh = np.ascontiguousarray(idx[:, :15])
    if use_ref:
        O.ref_grow(xs, nrm, neigh)
        O.ref_raster(xyz)
    else:
        O.grow(xs, nrm, neigh)
        img = O.raster(xs, mx[2] - mn[2], int(wh[0]), int(wh[1]))
        O.save_image(img)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as O

    O.orc()
    have_ref = O.ref() is not None
    cores = os.cpu_count() or 1
    n_full = args.points
    xyz_full = make_workload(args.workload, n_full)
    sample_n = args.ref_sample
    xyz = crop_sample(xyz_full, sample_n)
    del xyz_full
    times = []
    for it in range(args.warmup + args.steps):
        dt = cpu_path(xyz, have_ref)
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = len(xyz) * len(times) / total
    kind = "reference" if have_ref else "port"
    sample = (f"spatial crop of {len(xyz)} points of {args.workload} ({n_full} points); kNN/normals = oracle port "
              f"(Open3D 0.19 is not vendored), OpenMP x{cores}; grower+raster = "
              + ("the reference's own lines (oracle/_ref), single thread as in the reference"
                 if have_ref else "oracle port, single thread"))
    out = {
        "impl": "reference", "metric": "points/sec segmented end-to-end", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} synthetic, {n_full} points (CPU arm timed on a {len(xyz)}-point crop)"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from buildingsegment_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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

    n = args.points
    xyz = make_workload(args.workload, n, rank, world)
    x_lo, x_hi = int(rank * SLAB_M * 1000), int((rank + 1) * SLAB_M * 1000)
    if world > 1:  # this rank's slab of the city tile: x in [x_lo, x_hi) mm
        xyz = np.ascontiguousarray(xyz[(xyz[:, 0] >= x_lo) & (xyz[:, 0] < x_hi)])
    n = len(xyz)
    ctx = lib.Context(local)
    p = lib.default_params()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    d_xyz = torch.from_numpy(xyz).to(dev)
    W, H = None, None
    slab_info = {}
    if world > 1:
        from buildingsegment_b200 import slabs

        backend = slabs.CudaBackend(ctx, p)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        if world == 1:
            ctx.set_points_device(d_xyz.data_ptr(), n)
            ctx.run_device(p, lib.RUN_ALL)
            return None
        # tile origin -> halo exchange (NCCL P2P) -> kNN/normals + halo check -> grow -> cross-slab label merge
        r = slabs.segment_slab(backend, d_xyz, x_lo, x_hi, halo=args.halo)
        labels = r["labels"]
        # the tile's raster: this rank's pixel columns, bit-identical to the undivided tile's (slabs.raster_slab)
        img = slabs.raster_slab(backend, d_xyz, r["halo_l"], r["halo_r"], r["origin"], x_lo, x_hi, r["halo"], p.bin, p.bin_height)
        slab_info.update({k: r[k] for k in ("n_planes_total", "n_components", "halo", "n_halo")})
        slab_info.update({"raster_tile": [img["W"], img["H"]], "raster_columns": [img["x0"], img["x0"] + int(img["image"].shape[1])]})
        return labels

    # ---- device-resident leg ----
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ctx.reset_counters()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        t = ctx.timings()
        for k in ("bbox_keys", "sort", "cells", "knn", "knn_fallback", "grow", "finalize", "raster"):
            stage_ms[k] = stage_ms.get(k, 0.0) + t[k]
        last_t = t
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_dev = e0.elapsed_time(e1)
    launches = last_t["kernel_launches"]
    t_max = torch.tensor([ms_dev], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_dev = float(t_max.item())
    value = n * world * args.steps / (ms_dev * 1e-3)

    # ---- end-to-end leg: pinned host buffers through bseg_segment_host ----
    W, H = ctx.raster_size(p)
    h_xyz = torch.from_numpy(xyz).pin_memory()
    h_shift = torch.empty((n, 3), dtype=torch.int32).pin_memory()
    h_label = torch.empty(n, dtype=torch.int32).pin_memory()
    h_a = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    h_b = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
    np_xyz, np_shift, np_label, np_a, np_b = (h_xyz.numpy(), h_shift.numpy(), h_label.numpy(), h_a.numpy(), h_b.numpy())

    d_stage = torch.empty((n, 3), dtype=torch.int32, device=dev) if world > 1 else None
    png_host = [None]

    def step_host():
        if world == 1:
            return ctx.segment_host(p, np_xyz, np_shift, np_label, np_a, np_b)
        d_stage.copy_(h_xyz, non_blocking=True)  # H2D of the slab from pinned memory
        r = slabs.segment_slab(backend, d_stage, x_lo, x_hi, halo=args.halo)
        h_label.copy_(r["labels"].to(torch.int32))  # D2H of the canonical labels
        img = slabs.raster_slab(backend, d_stage, r["halo_l"], r["halo_r"], r["origin"], x_lo, x_hi, r["halo"], p.bin, p.bin_height)
        png_host[0] = (img["png_a"].cpu(), img["png_b"].cpu())  # D2H of this rank's columns of the two images
        torch.cuda.synchronize(dev)
        return r["n_planes_total"], 0, 0

    for _ in range(max(1, args.warmup // 2)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        npl, _, _ = step_host()
    torch.cuda.synchronize(dev)
    ms_e2e = (time.perf_counter() - t0) * 1e3
    t_max = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_max.item())
    e2e_value = n * world * args.steps / (ms_e2e * 1e-3)
    h2d = n * 12
    d2h = n * 12 + n * 4 + 2 * 3 * W * H if world == 1 else n * 4 + sum(int(t.numel()) for t in (png_host[0] or ()))

    if rank == 0:
        peak, peak_src = peaks()
        per_step = {k: v / args.steps for k, v in stage_ms.items()}
        dom = max(per_step, key=lambda k: per_step[k])
        stages = {}
        for k, ms in per_step.items():
            b = BYTES_PER_POINT.get(k)
            if b and ms > 0:
                gbs = b * n / (ms * 1e-3) / 1e9
                stages[k] = {"ms": round(ms, 3), "bytes_per_point": b, "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
            else:
                stages[k] = {"ms": round(ms, 3)}
        # dominant kernel: spec_grow_kernel (the grower's slices).  Algorithmic bytes of one Broad() call: the
        # node's row (4K) + per neighbour position 12, normal 24, state/label/reservation ~10; calls
        # per launch = (committed + released) calls / launches; duration = CUDA events around every slice.
        slice_ms = float(last_t["grow_slice_ms"])
        calls = int(last_t["grow_steps"]) - int(last_t["grow_tiny_tx"]) + int(last_t["grow_wasted_steps"])
        slices = max(1, int(last_t["grow_rounds"]) - 1)
        ach = BYTES_PER_POINT["grow"] * calls / (slice_ms * 1e-3) / 1e9 if slice_ms > 0 else None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                traffic = json.load(f).get("spec_grow_kernel_dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "spec_grow_kernel (plane grower slices; dependent-gather latency bound, "
                                              "see DESIGN.md 6)",
                    "achieved": round(ach, 2) if ach else None, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": round(ach / peak, 5) if ach else None, "traffic": traffic,
                    "launches": slices, "avg_launch_ms": round(slice_ms / slices, 3),
                    "bytes_per_call": BYTES_PER_POINT["grow"], "calls_per_launch": round(calls / slices, 1),
                    "dominant_stage": dom, "stages": stages}
        cpu = None
        if world == 1 and not args.no_cpu:
            import oracle_lib as O

            O.orc()
            sample = crop_sample(xyz, args.cpu_sample)
            dt = cpu_path(sample, False)
            cpu = {"value": len(sample) / dt, "unit": "points/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": f"spatial crop of {len(sample)} points of the same cloud, one pass, {dt:.1f} s; kNN/normals "
                             f"OpenMP on all cores, grower/raster single thread (as the reference)"}
        out = {
            "metric": "points/sec segmented end-to-end", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": f"{args.workload} synthetic suburban block, {n} points per GPU, reference defaults "
                                   f"(K=15, r=100 mm, max_nn=50, 300 mm, 0.88, 400)" if world == 1 else
                                   f"C5-style city tile, one {int(SLAB_M)} m C2-like x-slab of ~{n} points per GPU (tile origin, NCCL halo "
                                   f"exchange, halo check, per-slab segmentation, cross-slab label merge), reference defaults",
                       "points_per_gpu": n, "l2": "inputs larger than L2 (cloud + neighbour rows >> 126 MB)",
                       "grow_engine": "sweeper + speculative growers (grow_mode 0)", "planes": int(npl),
                       **({"slabs": slab_info} if world > 1 else {})},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "grow": {"steps": int(last_t["grow_steps"]), "rounds": int(last_t["grow_rounds"]),
                     "n_unresolved_knn": int(last_t["n_unresolved"]), "n_big_items": int(last_t["n_big_cells"]),
                     "wasted_steps": int(last_t["grow_wasted_steps"]), "sweep_iters": int(last_t["grow_sweep_iters"]),
                     "tiny_tx": int(last_t["grow_tiny_tx"]), "seq_fallbacks": int(last_t["grow_seq_fallbacks"]),
                     "head_steps": int(last_t["grow_head_steps"]), "head_ms": round(last_t["grow_head_ns"] / 1e6, 3),
                     "sweep_ms": round(last_t["grow_sweep_ns"] / 1e6, 3), "at_fails": int(last_t["grow_at_fails"])},
        }
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--cpu-sample", type=int, default=3_000_000)
    ap.add_argument("--ref-sample", type=int, default=60_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--halo", type=int, default=2000, help="initial slab halo width, mm (N > 1); doubled until sufficient")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
