#!/usr/bin/env python
"""Generates tests/golden/*.npz -- golden vectors of the reference's OWN code for this path.

Run in the build container (needs /root/reference): `python tests/golden/make_golden.py`.

What produces the expected outputs: oracle/_ref/libbseg_ref.so, i.e. the reference's own source lines
(my_function.h:25-30,89-123, my_function.cpp:180-275, TMC3.cpp:44-200) compiled where they lie by
oracle/build_ref.sh -- NOT the oracle restatement.  The vectors therefore pin the oracle port and the CUDA
path to the reference itself for the grower (a7-a11), set_plane_color (a10) and the raster (a3, a13-a15),
and they travel to the GPU box, where /root/reference does not exist.

The grower's inputs (shifted cloud, normals, neighbour rows) are stored too: the Open3D stage that
produces them in the reference is not vendored, so the fixture fixes them at the values the
oracle port computed when the fixture was made; tests re-derive them and require equality first.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402
import oracle_lib as O  # noqa: E402

# (fixture name, generator in tests/cases.py, kwargs)
GROW_CASES = [
    ("building_shuffled", "building", dict(n=30000, order="shuffled")),
    ("building_scan", "building", dict(n=30000, order="scan")),
    ("block", "block", dict(n=60000)),
    ("quantised", "quantised", dict(n=30000)),
    ("far_offset_wrap", "far_offset", dict()),
    ("grid_row", "grid_plane", dict(nx=80, ny=80, order="row")),
]
RASTER_CASES = [
    ("building_shuffled", "building", dict(n=30000, order="shuffled")),
    ("block", "block", dict(n=60000)),
    ("tiny200", "tiny", dict(n=200)),
]


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    assert O.ref() is not None, "oracle/_ref is not built (needs /root/reference): run oracle/build_ref.sh"
    for name, gen, kw in GROW_CASES:
        xyz = getattr(cases, gen)(**kw)
        P = O.pipeline(xyz)
        r = O.ref_grow(P["xyz"], P["normals"], P["neigh"])
        out = dict(
            xyz=P["xyz"], neigh_sha=sha(P["neigh"]), normals_sha=sha(P["normals"]),
            neigh_head=P["neigh"][:256], normals_head=P["normals"][:256],
            plane_idx=r.plane_idx, label=r.label, plane_seed=r.plane_seed, plane_normal=r.plane_normal,
            plane_center=r.plane_center, plane_off=r.plane_off, point_idx_sha=sha(r.point_idx),
            colors_sha=sha(r.colors), n_planes=np.int64(r.n_planes),
        )
        np.savez_compressed(os.path.join(HERE, f"grow_{name}.npz"), **out)
        print(name, "planes", r.n_planes, "points", len(xyz))
    for name, gen, kw in RASTER_CASES:
        xyz = getattr(cases, gen)(**kw)
        xs, W, H, img = O.ref_raster(xyz)
        # the three uint8 images the reference's own save_image lines hand to stbi_write_png, decoded back
        # from the files it wrote (GBK file names, TMC3.cpp:98,108,119)
        import tempfile

        import cv2

        out = dict(xyz_shifted=xs, W=np.int64(W), H=np.int64(H), image_ch0_sha=sha(img[..., 0]), image_ch2_sha=sha(img[..., 2]))
        if W * H <= 40000:  # the doubles themselves for the small images (channel 1 goes through std::log)
            out["image"] = img
        with tempfile.TemporaryDirectory() as d:
            O.ref_raster(xyz, d)
            for key, fn in (("png_height", "平均高度.png"), ("png_count", "像素数量.png"), ("png_both", "像素数量+高度.png")):
                raw = np.frombuffer(open(os.path.join(os.fsencode(d), fn.encode("gbk")), "rb").read(), np.uint8)
                out[key] = np.ascontiguousarray(cv2.imdecode(raw, cv2.IMREAD_COLOR)[..., ::-1])
        np.savez_compressed(os.path.join(HERE, f"raster_{name}.npz"), **out)
        print(name, W, H)


if __name__ == "__main__":
    main()
