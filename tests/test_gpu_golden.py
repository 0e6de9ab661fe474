"""GPU parity against the committed golden vectors of the reference's own lines (tests/golden/): the CUDA
path through the C ABI must reproduce planeIdx, labels, plane lists/models, colours and the PNG pixels of
the reference itself, bit for bit.  Needs neither /root/reference nor oracle/_ref."""
import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from buildingsegment_b200 import lib

    c = lib.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("name", G.names("grow"))
def test_cuda_grower_matches_reference_vectors(ctx, name, mode):
    from buildingsegment_b200 import lib

    g = G.load("grow", name)
    xyz = g["xyz"]
    p = lib.default_params(grow_mode=mode)
    ctx.set_points(xyz)
    neigh, nrm, _ = ctx.knn_normals(p)
    assert np.array_equal(G.sha(neigh), g["neigh_sha"]) and np.array_equal(G.sha(nrm), g["normals_sha"])
    pidx, label, npl = ctx.grow_planes(p)
    assert npl == int(g["n_planes"])
    assert np.array_equal(pidx, g["plane_idx"]) and np.array_equal(label, g["label"])
    seeds, normals, centers, off, idx = ctx.get_planes(npl)
    assert np.array_equal(seeds, g["plane_seed"]) and np.array_equal(off, g["plane_off"])
    assert np.array_equal(centers, g["plane_center"])
    assert np.array_equal(normals.view(np.int64), g["plane_normal"].view(np.int64))
    assert np.array_equal(G.sha(idx), g["point_idx_sha"])
    colors = ctx.paint(O.libc_plane_colors(npl))
    assert np.array_equal(G.sha(colors), g["colors_sha"])


@pytest.mark.parametrize("name", G.names("raster"))
def test_cuda_raster_matches_reference_vectors(ctx, name):
    from buildingsegment_b200 import lib

    g = G.load("raster", name)
    p = lib.default_params()
    mn, mx, xs = ctx.set_points(g["xyz_shifted"])
    assert np.array_equal(xs, g["xyz_shifted"])
    assert ctx.raster_size(p) == (int(g["W"]), int(g["H"]))
    img, a, b, c, th = ctx.raster(p)
    assert np.array_equal(G.sha(img[..., 0]), g["image_ch0_sha"]) and np.array_equal(G.sha(img[..., 2]), g["image_ch2_sha"])
    if "image" in g:  # all three channels of the reference's doubles, the libm log of channel 1 included
        assert np.array_equal(img.view(np.int64), g["image"].view(np.int64))
    assert np.array_equal(a, g["png_height"]) and np.array_equal(b, g["png_count"]) and np.array_equal(c, g["png_both"])
