"""The oracle port against the committed golden vectors of the reference's own lines (tests/golden/,
made by make_golden.py from oracle/_ref).  Runs without /root/reference and without a GPU."""
import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O


def test_fixtures_exist():
    assert len(G.names("grow")) >= 6 and len(G.names("raster")) >= 3


@pytest.mark.parametrize("name", G.names("grow"))
def test_oracle_grower_matches_reference_vectors(name):
    g = G.load("grow", name)
    xyz = g["xyz"]
    P = O.pipeline(xyz)  # already shifted: the shift is the identity
    assert np.array_equal(P["xyz"], xyz)
    # inputs of the grower are the ones the vectors were made with (Open3D stage restated)
    assert np.array_equal(G.sha(P["neigh"]), g["neigh_sha"]) and np.array_equal(G.sha(P["normals"]), g["normals_sha"])
    assert np.array_equal(P["neigh"][:256], g["neigh_head"])
    r = P["grow"]
    assert r.n_planes == int(g["n_planes"])
    assert np.array_equal(r.plane_idx, g["plane_idx"]) and np.array_equal(r.label, g["label"])
    assert np.array_equal(r.plane_seed, g["plane_seed"]) and np.array_equal(r.plane_off, g["plane_off"])
    assert np.array_equal(r.plane_center, g["plane_center"])
    assert np.array_equal(r.plane_normal.view(np.int64), g["plane_normal"].view(np.int64))
    assert np.array_equal(G.sha(r.point_idx), g["point_idx_sha"])
    colors = O.paint(len(xyz), r.plane_off, r.point_idx, O.libc_plane_colors(r.n_planes))
    assert np.array_equal(G.sha(colors), g["colors_sha"])


@pytest.mark.parametrize("name", G.names("raster"))
def test_oracle_raster_matches_reference_vectors(name):
    g = G.load("raster", name)
    xs = g["xyz_shifted"]
    x2, mn, mx, wh = O.bbox_shift(xs)
    W, H = int(g["W"]), int(g["H"])
    assert np.array_equal(x2, xs) and (int(wh[0]), int(wh[1])) == (W, H)
    img = O.raster(xs, mx[2] - mn[2], W, H)
    assert np.array_equal(G.sha(img[..., 0]), g["image_ch0_sha"]) and np.array_equal(G.sha(img[..., 2]), g["image_ch2_sha"])
    if "image" in g:
        ref = g["image"]
        # std::log (TMC3.cpp:161) is the platform libm's on both sides: the doubles are identical
        assert np.array_equal(img[..., 1].view(np.int64), ref[..., 1].view(np.int64))
    a, b, c, _ = O.save_image(img)
    assert np.array_equal(a, g["png_height"]) and np.array_equal(b, g["png_count"]) and np.array_equal(c, g["png_both"])
