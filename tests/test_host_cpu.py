"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/bseg.h
declares, fails loudly without a GPU, and the host-side ply::read / ply::write keep the reference's
value semantics (ply.cpp:88-504)."""
import os
import re
import subprocess

import numpy as np
import pytest

import plyio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "buildingsegment_b200", "host")


def test_libbseg_exports_every_declared_symbol():
    from buildingsegment_b200 import lib

    hdr = open(os.path.join(ROOT, "include", "bseg.h")).read()
    declared = set(re.findall(r"BSEG_API\s+[\w\s\*]+?\b(bseg_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(lib.EXPORTS)
    assert b"sm_100a" in L.bseg_version()


def test_no_cpu_fallback():
    import torch

    from buildingsegment_b200 import lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.BsegError) as e:
        lib.Context(0)
    assert e.value.code == -2  # BSEG_E_NODEVICE
    assert "no CPU path" in str(e.value)


def test_count_channel_is_the_platform_libm():
    """bseg_count_channel (the host half of the raster, TMC3.cpp:159-164): log(v + 1) of the platform's libm --
    math.log calls the same function -- plus the bias when non-zero; zero stays zero; the maximum is returned."""
    import math

    from buildingsegment_b200 import lib

    rng = np.random.default_rng(5)
    v = np.concatenate([np.zeros(100), rng.random(5000) * 4, rng.integers(1, 3000, 5000).astype(np.float64),
                        10.0 ** rng.uniform(-12, 6, 5000)])
    want = np.array([0.0 if x == 0 else (lambda y: y + 20.0 if y != 0 else y)(math.log(x + 1)) for x in v])
    got = v.copy()
    m = lib.count_channel(got, 20.0)
    assert np.array_equal(got.view(np.int64), want.view(np.int64))
    assert m == want.max()
    big = np.tile(v, 20)  # above the threading threshold
    m2 = lib.count_channel(big, 20.0)
    assert np.array_equal(big.view(np.int64), np.tile(want, 20).view(np.int64)) and m2 == m


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "buildingsegment_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and fn != "bseg_arith.h":
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle_lib" not in src and "libbseg_oracle" not in src and "oracle/" not in src, fn


@pytest.fixture(scope="module")
def ply_check(tmp_path_factory):
    out = tmp_path_factory.mktemp("bin") / "ply_check"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", HOST, os.path.join(ROOT, "tests", "host", "ply_check.cpp"),
                    os.path.join(HOST, "ply.cpp"), "-o", str(out)], check=True)
    return str(out)


def _run(ply_check, tmp_path, xyz, rgb, fmt, ptype, ascii_out=False):
    src = str(tmp_path / "in.ply")
    dst = str(tmp_path / "out.ply")
    dump = str(tmp_path / "dump.bin")
    plyio.write_ply_xyz_rgb(src, xyz, rgb, fmt, ptype)
    subprocess.run([ply_check, src, "1000", dst, dump] + (["a"] if ascii_out else []), check=True)
    raw = open(dump, "rb").read()
    n = int(np.frombuffer(raw[:8], np.int64)[0])
    hc = int(np.frombuffer(raw[8:12], np.int32)[0])
    pos = np.frombuffer(raw[12:12 + 12 * n], np.int32).reshape(n, 3)
    col = np.frombuffer(raw[12 + 12 * n:], np.uint16).reshape(n, 3) if hc else None
    return pos, col, dst


@pytest.mark.parametrize("fmt,ptype", [("binary_little_endian", "float"), ("binary_little_endian", "float64"),
                                       ("ascii", "float")])
def test_ply_read_semantics(ply_check, tmp_path, fmt, ptype):
    rng = np.random.default_rng(1)
    xyz = rng.uniform(-50, 50, (2000, 3))
    rgb = rng.integers(0, 256, (2000, 3))
    pos, col, dst = _run(ply_check, tmp_path, xyz, rgb, fmt, ptype)
    if ptype == "float" and fmt != "ascii":
        want = np.trunc(xyz.astype(np.float32).astype(np.float64) * 1000.0)
    else:
        want = np.trunc(xyz * 1000.0)
    assert np.array_equal(pos, want.astype(np.int32))  # int32(value * scale), truncation toward zero
    assert np.array_equal(col, rgb[:, [1, 2, 0]])       # stored [G, B, R] (ply.cpp:412-414)
    header, oxyz, ocol = plyio.read_ply_ref_output(dst)
    assert header[:2] == ["ply", "format binary_little_endian 1.0"]
    assert header[3:9] == ["property float64 x", "property float64 y", "property float64 z",
                           "property uchar green", "property uchar blue", "property uchar red"]
    assert header[-3:] == ["element face 0", "property list uint8 int32 vertex_index", "end_header"]
    assert np.array_equal(oxyz, pos.astype(np.float64)) and np.array_equal(ocol, col)


def test_ply_without_colours_and_ascii_out(ply_check, tmp_path):
    xyz = np.array([[0.0015, -0.0015, 1.9999], [3.0, 4.0, 5.0]])
    pos, col, dst = _run(ply_check, tmp_path, xyz, None, "binary_little_endian", "float64", ascii_out=True)
    assert col is None and pos.tolist() == [[1, -1, 1999], [3000, 4000, 5000]]
    lines = open(dst).read().splitlines()
    assert lines[1] == "format ascii 1.0" and lines[3] == "property float x"
    assert lines[-2:] == ["1.00000 -1.00000 1999.00000", "3000.00000 4000.00000 5000.00000"]


def test_synth_generators_are_deterministic():
    from buildingsegment_b200 import synth

    a = synth.make("C1", 5000)
    b = synth.make("C1", 5000)
    assert a.dtype == np.int32 and a.shape == (5000, 3) and np.array_equal(a, b)
    c2 = synth.make("C2", 20000)
    assert c2.shape == (20000, 3) and c2[:, 2].max() < 20000
    v = synth.make("C4", 3000, bits=6)
    assert len(np.unique(v, axis=0)) == len(v)  # voxelised: unique lattice points


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host cores alone (no GPU): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--points", "300000",
                        "--ref-sample", "6000", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config"):
        assert k in line


def test_bench_labels_checksum_is_split_invariant():
    """bench.py's labels checksum: the parts of contiguous chunks add up (as the int64 all_reduce at N > 1 adds them,
    wrapping mod 2^64) to the checksum of the undivided tile."""
    import importlib.util

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rng = np.random.default_rng(7)
    lab = rng.integers(0, 1 << 30, size=3_000_001).astype(np.int32)  # large labels: the weighted sum wraps many times
    whole = torch.tensor(bench.labels_checksum_parts(lab, 0), dtype=torch.int64)
    for cuts in ([0, 1_234_567, len(lab)], [0, 1, 2_000_000, 2_000_001, len(lab)]):
        acc = torch.zeros(2, dtype=torch.int64)
        for a, b in zip(cuts[:-1], cuts[1:]):
            acc += torch.tensor(bench.labels_checksum_parts(lab[a:b], a), dtype=torch.int64)  # int64 add wraps like the all_reduce
        assert torch.equal(acc, whole)
    moved = lab.copy()
    moved[[5, 6]] = moved[[6, 5]]
    if moved[5] != moved[6]:
        assert bench.labels_checksum_parts(moved, 0)[1] != bench.labels_checksum_parts(lab, 0)[1]  # position matters
