"""extracted_contour (my_function.cpp:8-145) without OpenCV, checked against cv2 (the test oracle; the product never
imports it): contours of bseg_find_contours == cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE / NONE) -- same
contours, same order, same first point and orientation; contourArea / arcLength; the 2-pixel overlay against
cv2.drawContours; on the GPU, threshold + close == cv2.threshold + cv2.morphologyEx."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from buildingsegment_b200 import lib  # noqa: E402


def _masks():
    rng = np.random.default_rng(7)
    out = []
    for t in range(120):
        H, W = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        dens = float(rng.choice([0.05, 0.2, 0.4, 0.55, 0.7, 0.9, 1.0]))
        m = (rng.random((H, W)) < dens).astype(np.uint8) * 255
        if t % 3 == 0 and H > 4 and W > 4:
            m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)), iterations=2)
        out.append(m)
    # nested: a ring with an island in its hole, twice (RETR_EXTERNAL must skip what lies inside a hole)
    m = np.zeros((40, 60), np.uint8)
    m[3:30, 4:40] = 255
    m[8:25, 9:35] = 0
    m[12:20, 14:28] = 255
    m[14:18, 17:25] = 0
    m[15:17, 19:22] = 255
    m[33:38, 50:58] = 255
    out.append(m)
    out.append(np.zeros((5, 5), np.uint8))
    out.append(np.full((6, 9), 255, np.uint8))
    return out


@pytest.mark.parametrize("simple", [True, False])
def test_contours_match_opencv(simple):
    mode = cv2.CHAIN_APPROX_SIMPLE if simple else cv2.CHAIN_APPROX_NONE
    n_contours = 0
    for m in _masks():
        ref, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, mode)
        got = lib.find_contours(m, simple)
        assert len(got) == len(ref)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r.reshape(-1, 2))
            a, l = lib.contour_measure(g)
            assert a == cv2.contourArea(r)
            assert abs(l - cv2.arcLength(r, True)) <= 1e-4 * max(1.0, l)
        n_contours += len(ref)
    assert n_contours > 500


def test_overlay_is_close_to_drawcontours():
    rng = np.random.default_rng(3)
    m = np.zeros((300, 400), np.uint8)
    for _ in range(12):
        x, y = int(rng.integers(20, 340)), int(rng.integers(20, 240))
        cv2.ellipse(m, (x, y), (int(rng.integers(8, 40)), int(rng.integers(8, 40))), float(rng.uniform(0, 180)), 0, 360, 255, -1)
    cs, _ = cv2.findContours(m.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    ref = np.zeros((300, 400, 3), np.uint8)
    for i in range(len(cs)):
        cv2.drawContours(ref, cs, i, (255, 255, 0), 2)
    got = np.zeros((300, 400, 3), np.uint8)
    for c in lib.find_contours(m, True):
        lib.draw_contour(got, c, (255, 255, 0))
    # OpenCV's fixed-point polygon fill is restated as a rule (contour.cu: contour_draw), not transcribed: identical on
    # these shapes, but the claim tested is "at most 1 % of the stroke pixels differ"
    a, b = ref.any(axis=2), got.any(axis=2)
    assert (a ^ b).sum() <= 0.01 * (a | b).sum(), ((a ^ b).sum(), (a | b).sum())
    assert set(map(tuple, got[b])) == {(255, 255, 0)}


@pytest.mark.gpu
def test_mask_on_the_device_matches_opencv():
    rng = np.random.default_rng(5)
    ctx = lib.Context(0)
    try:
        for (H, W) in [(1, 1), (7, 300), (257, 129), (1500, 2100)]:
            img = np.zeros((H, W, 3), np.uint8)
            occ = rng.random((H, W)) < 0.3
            img[..., 1] = np.where(occ, rng.integers(0, 256, (H, W)), rng.integers(0, 12, (H, W)))
            img[..., 0] = rng.integers(0, 256, (H, W))
            got = ctx.contour_mask(img, channel=1, thresh=10, iterations=2)
            _, th = cv2.threshold(np.ascontiguousarray(img[..., 1]), 10, 255, cv2.THRESH_BINARY)
            ref = cv2.morphologyEx(th, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)), anchor=(-1, -1),
                                   iterations=2)
            assert np.array_equal(got, ref)
    finally:
        ctx.close()
