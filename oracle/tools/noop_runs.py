#!/usr/bin/env python
"""Run-length statistics of Broad() calls that accept nothing, on a crop of the C2 cloud (analysis tool, CPU only).

  python oracle/tools/noop_runs.py [points_in_crop]

Prints the share of calls that accept nothing, the run-length distribution of such calls for the largest plane, and
what a pair-at-a-time engine vs a 32/64-node skip batch would need in dependent steps.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import bench  # noqa: E402
import oracle_lib as O  # noqa: E402
from buildingsegment_b200 import synth  # noqa: E402


def main():
    target = int(sys.argv[1]) if len(sys.argv) > 1 else 2_500_000
    so = os.path.join(HERE, "_noop_runs.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fwrapv", "-shared", "-fPIC", "-o", so,
                    os.path.join(HERE, "noop_runs.c"), "-lm"], check=True)
    L = C.CDLL(so)
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
    i64p = np.ctypeslib.ndpointer(np.int64, flags="C")
    L.sim_grow.argtypes = [i32p, C.c_int64, f64p, i32p, C.c_int, C.c_int, C.c_int, C.c_double, u8p, C.c_int64, i64p,
                           C.c_int64, i64p]
    L.sim_grow.restype = C.c_int64
    xyz = bench.crop_sample(synth.make("C2", 10_000_000), target)
    xs, mn, mx, wh = O.bbox_shift(xyz)
    idx, d2 = O.knn(xs, 50, cell=100)
    nrm, _, _ = O.normals(xs, idx, d2, 100.0, 50)
    neigh = np.ascontiguousarray(idx[:, :15])
    cap = len(xs) * 2
    cnts = np.zeros(cap, np.uint8)
    txo = np.zeros(100000, np.int64)
    ntx = np.zeros(1, np.int64)
    nc = L.sim_grow(xs, len(xs), nrm, neigh, 15, 300, 400, 0.88, cnts, cap, txo, len(txo) - 1, ntx)
    cn = cnts[:nc]
    ntx = int(ntx[0])
    txo = txo[: ntx + 1]
    acc = cn & 0x3F
    z = acc == 0
    print(f"{len(xs)} points, {ntx} transactions past depth 0, {nc} calls, {100 * z.mean():.1f} % accept nothing; "
          f"{100 * (z & ((cn & 0x80) != 0)).sum() / max(1, z.sum()):.1f} % of those still see a free neighbour")
    big = int(np.argmax(np.diff(txo)))
    c = acc[txo[big]: txo[big + 1]]
    zz = (c == 0).astype(np.int8)
    d = np.diff(np.concatenate([[0], zz, [0]]))
    rl = np.nonzero(d == -1)[0] - np.nonzero(d == 1)[0]
    print(f"largest plane: {len(c)} calls, {100 * zz.mean():.1f} % accept nothing, {len(rl)} runs, mean {rl.mean():.1f}, "
          f"max {rl.max()}; in runs >= 8: {100 * rl[rl >= 8].sum() / zz.sum():.1f} %, >= 32: {100 * rl[rl >= 32].sum() / zz.sum():.1f} %")


if __name__ == "__main__":
    main()
