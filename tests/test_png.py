"""The PNG writer (include/bseg.h bseg_png_*; host-only, no GPU): byte-identical to stbi_write_png of the reference's
vendored stb_image_write v1.16 (TMC3.cpp:98,108,119) -- against the reference's code when it is present, against the
committed digests made from it everywhere, decodable by an independent decoder (cv2), asynchronous writes included."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as O
import png_cases

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "png_stb.json")))


@pytest.mark.parametrize("name,img", list(png_cases.images()), ids=[n for n, _ in png_cases.images()])
def test_png_bytes_match_stb_digests(name, img):
    from buildingsegment_b200 import lib

    png = lib.png_encode(img)
    assert len(png) == GOLD[name]["bytes"]
    assert hashlib.sha256(png).hexdigest() == GOLD[name]["sha256"]


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref (the reference's own stb_image_write) is not built here")
def test_png_bytes_match_the_reference_encoder():
    from buildingsegment_b200 import lib

    rng = np.random.default_rng(7)
    for name, img in png_cases.images():
        assert lib.png_encode(img) == O.ref_png(img), name
    for _ in range(40):  # random shapes / sparsities / component counts
        h, w, c = int(rng.integers(1, 90)), int(rng.integers(1, 130)), int(rng.integers(1, 5))
        img = (rng.integers(0, 256, (h, w, c)) * (rng.random((h, w, c)) < rng.random())).astype(np.uint8)
        assert lib.png_encode(img) == O.ref_png(img), (h, w, c)


def test_png_decodes_and_async_writes(tmp_path):
    import ctypes as C

    import cv2

    from buildingsegment_b200 import lib

    L = lib.lib()
    paths = []
    for name, img in png_cases.images():
        if img.ndim != 3 or img.shape[2] != 3:
            continue
        p = str(tmp_path / f"{name}.png")
        assert L.bseg_png_write_async(p.encode(), img.ctypes.data, img.shape[1], img.shape[0], 3, 0) == 0
        paths.append((p, img.copy()))
        img[:] = 0  # the worker owns a copy: the caller's buffer may change at once
    assert L.bseg_png_wait() == 0
    for p, img in paths:
        raw = np.frombuffer(open(p, "rb").read(), np.uint8)
        assert raw.tobytes() == lib.png_encode(img)
        dec = cv2.imdecode(raw, cv2.IMREAD_COLOR)[..., ::-1]
        assert np.array_equal(dec, img), p
    assert L.bseg_png_write_async(b"/nonexistent_dir/x.png", paths[0][1].ctypes.data, 4, 4, 3, 0) == 0
    assert L.bseg_png_wait() != 0  # the failure surfaces at the join
