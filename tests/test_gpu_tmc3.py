"""End-to-end through the drop-in driver: PLY -> ./tmc3_b200 (reference CLI, TMC3.cpp:202-229) -> PLY,
compared with the oracle pipeline: shifted millimetre coordinates, [G,B,R] plane colours drawn from
libc rand() exactly like set_plane_color, and the three raster PNGs."""
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle_lib as O
import plyio

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tmc3_driver_matches_oracle(tmp_path):
    import cv2

    exe = os.path.join(ROOT, "tmc3_b200")
    assert os.path.exists(exe), "tmc3_b200 not built (run __graft_entry__.build())"
    mm = cases.building(50000)
    xyz_m = mm.astype(np.float64) / 1000.0 + 0.0004  # metres; +0.4 mm keeps float32 truncation away from .0
    rgb = np.random.default_rng(0).integers(0, 256, (len(mm), 3))
    src, dst = str(tmp_path / "in.ply"), str(tmp_path / "out.ply")
    plyio.write_ply_xyz_rgb(src, xyz_m, rgb)
    rdir = str(tmp_path) + "/"
    r = subprocess.run([exe, f"-a={src}", f"-s={dst}", f"--raster={rdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # what ply::read produces from the file
    pts = np.trunc(xyz_m.astype(np.float32).astype(np.float64) * 1000.0).astype(np.int32)
    P = O.pipeline(pts)
    header, oxyz, ocol = plyio.read_ply_ref_output(dst)
    assert np.array_equal(oxyz, P["xyz"].astype(np.float64))  # main writes the SHIFTED mm coordinates (TMC3.cpp:221)
    g = P["grow"]
    want = O.paint(len(pts), g.plane_off, g.point_idx, O.libc_plane_colors(g.n_planes, seed=1))
    assert g.n_planes > 0 and np.array_equal(ocol, want.astype(np.uint8))
    W, H = int(P["wh"][0]), int(P["wh"][1])
    oa, ob, oc, _ = O.save_image(O.raster(P["xyz"], P["mx"][2] - P["mn"][2], W, H))
    for name, mine in (("平均高度.png", oa), ("像素数量.png", ob), ("像素数量+高度.png", oc)):
        raw = np.frombuffer(open(os.path.join(os.fsencode(rdir), name.encode("gbk")), "rb").read(), np.uint8)
        png = cv2.imdecode(raw, cv2.IMREAD_COLOR)[..., ::-1]
        assert np.array_equal(png, mine), name
    # the label image (extension): plane label of the highest point per pixel in the planes' colours
    raw = np.frombuffer(open(os.path.join(rdir, "labels.png"), "rb").read(), np.uint8)
    lab_png = cv2.imdecode(raw, cv2.IMREAD_COLOR)[..., ::-1]
    _, olab = O.label_raster(P["xyz"], g.label, W, H, O.libc_plane_colors(g.n_planes, seed=1))
    assert np.array_equal(lab_png, olab)


def test_tmc3_driver_reports_missing_input(tmp_path):
    exe = os.path.join(ROOT, "tmc3_b200")
    r = subprocess.run([exe, "-a=/nonexistent.ply", f"-s={tmp_path}/o.ply"], capture_output=True, text=True)
    assert r.returncode != 0
