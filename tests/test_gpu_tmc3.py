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


def _ref_obj(contours, cols, rows):
    """csa.obj as my_function.cpp:64-128 writes it (ofstream << float: 6 significant digits)."""
    out = [b"# \xb4\xd3\xc2\xd6\xc0\xaa\xc9\xfa\xb3\xc9\xb5\xc43D\xc4\xa3\xd0\xcd\n",
           b"# \xc2\xd6\xc0\xaa\xca\xfd\xc1\xbf: " + str(len(contours)).encode() + b"\n",
           b"# \xb6\xa5\xb5\xe3\xb9\xe9\xd2\xbb\xbb\xaf\xb5\xbd\xb7\xb6\xce\xa7 [0,1] (x,y)\n\n"]
    groups, v = [], 1
    for c in contours:
        g = []
        for x, y in c.reshape(-1, 2):
            fx = np.float32(x) / np.float32(cols)
            fy = np.float32(1.0) - np.float32(y) / np.float32(rows)
            out.append(("v %g %g 0.0\n" % (fx, fy)).encode())
            out.append(("v %g %g 1\n" % (fx, fy)).encode())
            g += [v, v + 1]
            v += 2
        groups.append(g)
    out.append(b"\n# \xb2\xe0\xc3\xe6 (\xcb\xc4\xb1\xdf\xd0\xce\xc3\xe6)\n")
    for g in groups:
        n = len(g) // 2
        for i in range(n):
            j = (i + 1) % n
            out.append(("f %d %d %d %d\n" % (g[2 * i], g[2 * j], g[2 * j + 1], g[2 * i + 1])).encode())
    return b"".join(out)


def test_tmc3_driver_contours_and_classes(tmp_path):
    """--contours: extracted_contour (my_function.cpp:8-145) on the count image the driver has just written, against the
    same steps done with cv2: overlay pixels, its flip, and csa.obj byte for byte.  --classes: the class line."""
    import cv2

    exe = os.path.join(ROOT, "tmc3_b200")
    mm = cases.block(n=400000)
    xyz_m = mm.astype(np.float64) / 1000.0 + 0.0004
    rgb = np.zeros((len(mm), 3), np.int64)
    src, dst = str(tmp_path / "in.ply"), str(tmp_path / "out.ply")
    plyio.write_ply_xyz_rgb(src, xyz_m, rgb)
    rdir = str(tmp_path) + "/"
    r = subprocess.run([exe, f"-a={src}", f"-s={dst}", f"--raster={rdir}", f"--contours={rdir}", "--classes"], capture_output=True,
                       text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "roof" in r.stdout and "facade" in r.stdout
    raw = np.frombuffer(open(os.path.join(os.fsencode(rdir), "像素数量.png".encode("gbk")), "rb").read(), np.uint8)
    img = cv2.imdecode(raw, cv2.IMREAD_COLOR)  # B,G,R like the reference's imread
    _, th = cv2.threshold(np.ascontiguousarray(img[..., 1]), 10, 255, cv2.THRESH_BINARY)
    morphed = cv2.morphologyEx(th, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)), anchor=(-1, -1), iterations=2)
    contours, _ = cv2.findContours(morphed, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    assert len(contours) > 0
    result = img.copy()
    kept = [c for c in contours if cv2.contourArea(c) > 500 and cv2.arcLength(c, True) > 100]
    for i in range(len(kept)):
        cv2.drawContours(result, kept, i, (255, 255, 0), 2)
    got = cv2.imdecode(np.frombuffer(open(rdir + "extracted_contours.png", "rb").read(), np.uint8), cv2.IMREAD_COLOR)
    flip = cv2.imdecode(np.frombuffer(open(rdir + "extracted_contours_flip.png", "rb").read(), np.uint8), cv2.IMREAD_COLOR)
    assert len(kept) > 0
    assert (got != result).any(axis=2).sum() <= 0.01 * (result != img).any(axis=2).sum()  # (stroke rule, see test_contour.py)
    assert np.array_equal(flip, got[::-1])
    assert open(os.path.join(str(tmp_path), "csa.obj"), "rb").read() == _ref_obj(contours, img.shape[1], img.shape[0])
