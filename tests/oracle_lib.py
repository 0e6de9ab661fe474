"""ctypes bindings for the CPU oracle (oracle/libbseg_oracle.so) and, when built, the reference's
own lines (oracle/_ref/libbseg_ref.so).  TEST INFRASTRUCTURE -- imported by tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs only; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_ORC = None
_REF = None

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def orc():
    global _ORC
    if _ORC is None:
        path = os.path.join(ORACLE_DIR, "libbseg_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_bbox_shift.argtypes = [i32p, C.c_int64, i32p, i32p, C.c_int, i32p]
        L.orc_knn.argtypes = [i32p, C.c_int64, C.c_int, C.c_int32, i32p, C.c_void_p]
        L.orc_normals.argtypes = [i32p, C.c_int64, i32p, i64p, C.c_int, C.c_double, C.c_int, f64p, C.c_void_p, C.c_void_p]
        L.orc_grow.argtypes = [i32p, C.c_int64, f64p, i32p, C.c_int, C.c_int, C.c_int, C.c_double, i32p, i32p,
                               i32p, f64p, i32p, i64p, i32p, C.c_int64, C.c_int64, i64p]
        L.orc_grow.restype = C.c_int64
        L.orc_paint.argtypes = [C.c_int64, C.c_int64, i64p, i32p, u16p, u16p]
        L.orc_ground_th.argtypes = [i32p, C.c_int64, C.c_int32, C.c_int]
        L.orc_ground_th.restype = C.c_double
        L.orc_raster.argtypes = [i32p, C.c_int64, C.c_int32, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, f64p]
        L.orc_raster_th.argtypes = [i32p, C.c_int64, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, f64p]
        L.orc_raster_th_sums.argtypes = [i32p, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_int, f64p]
        L.orc_save_image.argtypes = [f64p, C.c_int, C.c_int, u8p, u8p, u8p, f64p]
        for f in ("orc_acos", "orc_cos", "orc_log"):
            getattr(L, f).argtypes = [C.c_double]
            getattr(L, f).restype = C.c_double
        L.orc_eigen.argtypes = [f64p, f64p, f64p]
        _ORC = L
    return _ORC


def ref():
    """The reference's own grower/raster lines, or None when oracle/_ref was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libbseg_ref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/tmc3"):
                build()
            if not os.path.exists(path):
                return None
        L = C.CDLL(path)
        L.ref_grow.argtypes = [i32p, C.c_int64, f64p, i32p, C.c_int, i32p, i32p, i32p, f64p, i32p, i64p, i32p,
                               C.c_int64, C.c_int64, u16p]
        L.ref_grow.restype = C.c_int64
        L.ref_raster.argtypes = [i32p, C.c_int64, i32p, f64p, C.c_int64, C.c_char_p]
        L.ref_png_encode.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int64]
        L.ref_png_encode.restype = C.c_int64
        _REF = L
    return _REF


# ---------------------------------------------------------------------------------------------
DEFAULTS = dict(K=15, radius=100.0, max_nn=50, th_thickness=300, th_point_count=400, th_dot=0.88,
                bin=100, bin_height=1000, count_bias=20.0, grow_radius=0.0)


def bbox_shift(xyz, bin=100):
    xyz = np.ascontiguousarray(xyz, np.int32).copy()
    mn = np.zeros(3, np.int32)
    mx = np.zeros(3, np.int32)
    wh = np.zeros(2, np.int32)
    orc().orc_bbox_shift(xyz, len(xyz), mn, mx, bin, wh)
    return xyz, mn, mx, wh


def knn(xyz, kq, cell=100):
    n = len(xyz)
    idx = np.empty((n, kq), np.int32)
    d2 = np.empty((n, kq), np.int64)
    rc = orc().orc_knn(np.ascontiguousarray(xyz, np.int32), n, kq, cell, idx, d2.ctypes.data)
    assert rc == 0, rc
    return idx, d2


def normals(xyz, knn_idx, knn_d2, radius=100.0, max_nn=50):
    n, kq = knn_idx.shape
    out = np.empty((n, 3), np.float64)
    curv = np.empty(n, np.float64)
    nh = np.empty(n, np.int32)
    rc = orc().orc_normals(np.ascontiguousarray(xyz, np.int32), n, knn_idx, knn_d2, kq, radius, max_nn, out,
                           curv.ctypes.data, nh.ctypes.data)
    assert rc == 0, rc
    return out, curv, nh


class GrowResult:
    pass


def _grow_common(fn, xyz, nrm, neigh, K, extra_args, with_colors):
    n = len(xyz)
    r = GrowResult()
    r.plane_idx = np.empty(n, np.int32)
    r.label = np.empty(n, np.int32)
    cap_p = n // 100 + 16
    cap_i = 2 * n + 16
    r.plane_seed = np.empty(cap_p, np.int32)
    r.plane_normal = np.empty((cap_p, 3), np.float64)
    r.plane_center = np.empty((cap_p, 3), np.int32)
    r.plane_off = np.zeros(cap_p + 1, np.int64)
    r.point_idx = np.empty(cap_i, np.int32)
    return r, cap_p, cap_i


def grow(xyz, nrm, neigh, K=15, th_thickness=300, th_point_count=400, th_dot=0.88):
    xyz = np.ascontiguousarray(xyz, np.int32)
    nrm = np.ascontiguousarray(nrm, np.float64)
    neigh = np.ascontiguousarray(neigh, np.int32)
    r, cap_p, cap_i = _grow_common(None, xyz, nrm, neigh, K, None, False)
    steps = np.zeros(1, np.int64)
    npl = orc().orc_grow(xyz, len(xyz), nrm, neigh, K, th_thickness, th_point_count, th_dot, r.plane_idx, r.label,
                         r.plane_seed, r.plane_normal, r.plane_center, r.plane_off, r.point_idx, cap_p, cap_i, steps)
    assert npl >= 0, npl
    _trim(r, npl)
    r.steps = int(steps[0])
    return r


def ref_grow(xyz, nrm, neigh, K=15):
    L = ref()
    assert L is not None
    xyz = np.ascontiguousarray(xyz, np.int32)
    nrm = np.ascontiguousarray(nrm, np.float64)
    neigh = np.ascontiguousarray(neigh, np.int32)
    r, cap_p, cap_i = _grow_common(None, xyz, nrm, neigh, K, None, True)
    r.colors = np.zeros((len(xyz), 3), np.uint16)
    npl = L.ref_grow(xyz, len(xyz), nrm, neigh, K, r.plane_idx, r.label, r.plane_seed, r.plane_normal,
                     r.plane_center, r.plane_off, r.point_idx, cap_p, cap_i, r.colors)
    assert npl >= 0, npl
    _trim(r, npl)
    return r


def _trim(r, npl):
    r.n_planes = int(npl)
    r.plane_seed = r.plane_seed[:npl]
    r.plane_normal = r.plane_normal[:npl]
    r.plane_center = r.plane_center[:npl]
    r.plane_off = r.plane_off[: npl + 1]
    r.point_idx = r.point_idx[: int(r.plane_off[npl])]


def libc_plane_colors(n_planes, seed=1):
    """The colour sequence set_plane_color draws: 55+rand()%200, three per plane, left to right."""
    libc = C.CDLL("libc.so.6")
    libc.srand(seed)
    out = np.empty((n_planes, 3), np.uint16)
    for p in range(n_planes):
        for k in range(3):
            out[p, k] = 55 + libc.rand() % 200
    return out


def paint(n, plane_off, point_idx, plane_rgb):
    colors = np.empty((n, 3), np.uint16)
    orc().orc_paint(n, len(plane_off) - 1, np.ascontiguousarray(plane_off, np.int64),
                    np.ascontiguousarray(point_idx, np.int32), np.ascontiguousarray(plane_rgb, np.uint16), colors)
    return colors


def raster(xyz_shifted, zext, W, H, bin=100, bin_height=1000, bias=20.0):
    img = np.empty((H, W, 3), np.float64)
    orc().orc_raster(np.ascontiguousarray(xyz_shifted, np.int32), len(xyz_shifted), int(zext), bin, bin_height, bias,
                     int(W), int(H), img.reshape(-1))
    return img


def raster_th(xyz_shifted, th, W, H, bin=100, bias=20.0):
    img = np.empty((H, W, 3), np.float64)
    orc().orc_raster_th(np.ascontiguousarray(xyz_shifted, np.int32), len(xyz_shifted), float(th), bin, bias, int(W), int(H),
                        img.reshape(-1))
    return img


def raster_th_sums(xyz_shifted, th, W, H, bin=100):
    """The raster with channel 1 left as the weight sums (what bseg_raster_device leaves on the device)."""
    img = np.empty((H, W, 3), np.float64)
    orc().orc_raster_th_sums(np.ascontiguousarray(xyz_shifted, np.int32), len(xyz_shifted), float(th), bin, int(W), int(H),
                             img.reshape(-1))
    return img


def save_image(img):
    H, W, _ = img.shape
    a = np.empty((H, W, 3), np.uint8)
    b = np.empty((H, W, 3), np.uint8)
    c = np.empty((H, W, 3), np.uint8)
    mx = np.zeros(3, np.float64)
    orc().orc_save_image(np.ascontiguousarray(img).reshape(-1), W, H, a.reshape(-1), b.reshape(-1), c.reshape(-1), mx)
    return a, b, c, mx


def ref_raster(xyz_unshifted, out_dir=None):
    L = ref()
    assert L is not None
    xyz = np.ascontiguousarray(xyz_unshifted, np.int32).copy()
    wh = np.zeros(2, np.int32)
    ext = xyz.max(0).astype(np.int64) - xyz.min(0)
    cap = int((ext[0] // 100 + 2) * (ext[1] // 100 + 2) * 3)
    img = np.empty(cap, np.float64)
    d = None if out_dir is None else (out_dir.rstrip("/") + "/").encode()
    rc = L.ref_raster(xyz, len(xyz), wh, img, cap, d)
    assert rc == 0, rc
    return xyz, int(wh[0]), int(wh[1]), img.reshape(int(wh[1]), int(wh[0]), 3)


def ref_png(img):
    """stbi_write_png_to_mem of the reference's vendored stb on a uint8 [H][W][comp] image."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    comp = 1 if img.ndim == 2 else img.shape[2]
    cap = img.size * 2 + 4096
    out = np.empty(cap, np.uint8)
    n = ref().ref_png_encode(img.reshape(-1), w, h, comp, 0, out, cap)
    assert n > 0, n
    return out[:n].tobytes()


def pipeline(xyz_unshifted, **kw):
    """Whole reference path on the CPU oracle: shift -> kNN -> normals -> grow.  Returns a dict."""
    p = dict(DEFAULTS)
    p.update(kw)
    xyz, mn, mx, wh = bbox_shift(xyz_unshifted, p["bin"])
    kq = max(p["max_nn"], p["K"])
    idx, d2 = knn(xyz, kq, cell=max(1, int(p["radius"])))
    nrm, curv, nh = normals(xyz, idx, d2, p["radius"], p["max_nn"])
    neigh = np.ascontiguousarray(idx[:, : p["K"]])
    grow_rows = neigh
    if p["grow_radius"] > 0:  # "radius-search growing" (BASELINE config C2): entries beyond the radius do not exist
        cut = (d2[:, : p["K"]].astype(np.float64) >= p["grow_radius"] ** 2)
        cut[:, 0] = False
        grow_rows = np.ascontiguousarray(np.where(cut, -1, neigh).astype(np.int32))
    g = grow(xyz, nrm, grow_rows, p["K"], p["th_thickness"], p["th_point_count"], p["th_dot"])
    return dict(xyz=xyz, mn=mn, mx=mx, wh=wh, knn=idx, d2=d2, normals=nrm, curvature=curv, n_hyb=nh, neigh=neigh, grow=g)


def label_raster(xyz_shifted, label, W, H, plane_rgb=None, bin=100):
    """Label raster restated in numpy (include/bseg.h bseg_label_raster; no reference counterpart): per pixel the
    label of the highest point, ties to the lower index; the colour image from the set_plane_color sequence."""
    xs = np.asarray(xyz_shifted, np.int64)
    px = (xs[:, 1] // bin) * W + (xs[:, 0] // bin)
    order = np.lexsort((np.arange(len(xs)), -xs[:, 2], px))  # pixel, then z descending, then index ascending
    first = np.ones(len(xs), bool)
    first[1:] = px[order][1:] != px[order][:-1]
    win = order[first]
    lab = np.zeros(W * H, np.int32)
    lab[px[win]] = np.asarray(label, np.int32)[win]
    rgb = np.zeros((W * H, 3), np.uint8)
    if plane_rgb is not None and len(plane_rgb):
        m = lab > 0
        rgb[m] = np.asarray(plane_rgb, np.uint16)[lab[m] - 1].astype(np.uint8)
    return lab.reshape(H, W), rgb.reshape(H, W, 3)


def plane_classes(normals, centers, label, ground_z, facade_max_nz=0.3, roof_min_nz=0.7):
    """numpy restatement of bseg_plane_classes (include/bseg.h): DetectedPlane equations (my_function.h:41-46) and
    roof / facade / ground classes.  No reference behaviour exists for this (the record is never filled there)."""
    n = np.asarray(normals, np.float64).reshape(-1, 3)
    c = np.asarray(centers, np.float64).reshape(-1, 3)
    d = -((n[:, 0] * c[:, 0] + n[:, 1] * c[:, 1]) + n[:, 2] * c[:, 2])
    eq = np.concatenate([n, d[:, None]], axis=1)
    az = np.abs(n[:, 2])
    cls = np.full(len(n), 4, np.uint8)
    cls[az <= facade_max_nz] = 2
    hor = az >= roof_min_nz
    cls[hor & (c[:, 2] >= ground_z)] = 1
    cls[hor & (c[:, 2] < ground_z)] = 3
    lab = np.asarray(label)
    pt = np.where(lab > 0, np.concatenate([[0], cls])[np.clip(lab, 0, None)], 0).astype(np.uint8)
    return eq, cls, pt
