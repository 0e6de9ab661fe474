"""ctypes binding of libbseg.so (include/bseg.h) -- the C ABI a maintainer binds from the reference's
C++ (INTEGRATION.md shows the C++ side).  Used by tests/, bench.py and __graft_entry__.py.

There is no CPU path: if libbseg.so is missing or no CUDA device is usable, every entry fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSEG_LIB_PATH") or os.path.join(HERE, "libbseg.so")  # (override: A/B runs of two builds)
CSRC = os.path.join(HERE, "csrc")

EXPORTS = [
    "bseg_create", "bseg_destroy", "bseg_default_params", "bseg_last_error", "bseg_version",
    "bseg_set_points", "bseg_set_points_device", "bseg_knn_normals", "bseg_override_neigh_normals",
    "bseg_grow_planes", "bseg_set_grow_offset", "bseg_knn_device_results", "bseg_import_neigh_normals_device", "bseg_get_planes", "bseg_paint", "bseg_plane_classes", "bseg_contour_mask", "bseg_find_contours", "bseg_contour_measure", "bseg_draw_contour", "bseg_raster_size", "bseg_raster", "bseg_count_channel", "bseg_png_encode", "bseg_png_write", "bseg_png_write_async", "bseg_png_wait", "bseg_raster_device", "bseg_label_raster",
    "bseg_run_device", "bseg_segment_host", "bseg_get_timings", "bseg_reset_counters", "bseg_stream",
    "bseg_point_count", "bseg_plane_count", "bseg_set_owned", "bseg_set_origin", "bseg_device_results", "bseg_halo_check", "bseg_debug_sort_pairs", "bseg_debug_exclusive_scan",
]


class Params(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("max_nn", C.c_int32), ("radius", C.c_double),
        ("th_thickness", C.c_int32), ("th_point_count", C.c_int32), ("th_dot", C.c_double),
        ("bin", C.c_int32), ("bin_height", C.c_int32), ("count_bias", C.c_double),
        ("cell", C.c_int32), ("grow_mode", C.c_int32), ("grow_radius", C.c_double), ("reserved", C.c_int32 * 3),
    ]


class Timings(C.Structure):
    _fields_ = [(k, C.c_float) for k in
                ("h2d", "bbox_keys", "sort", "cells", "knn", "knn_fallback", "normals", "grow", "finalize",
                 "raster", "d2h", "total", "grow_slice_ms", "grow_sweep_ms", "raster_host")] + \
               [(k, C.c_int64) for k in ("n_unresolved", "grow_steps", "grow_rounds", "kernel_launches",
                                         "n_big_cells", "grow_wasted_steps", "grow_sweep_iters",
                                         "grow_tiny_tx", "grow_seq_fallbacks", "grow_head_steps", "grow_head_ns",
                                         "grow_sweep_ns", "grow_at_fails")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BsegError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libbseg error {code}: {msg}")
        self.code = code


_LIB = None


def build(force=False):
    """Compile libbseg.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-s", "-C", CSRC, "clean"], check=True)
    subprocess.run(["make", "-s", "-j8", "-C", CSRC], check=True)


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise BsegError(-2, f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32p, i64, i32 = C.c_void_p, C.c_void_p, C.c_int64, C.c_int32
        L.bseg_create.argtypes = [C.POINTER(vp), C.c_int]
        L.bseg_destroy.argtypes = [vp]
        L.bseg_destroy.restype = None
        L.bseg_default_params.argtypes = [C.POINTER(Params)]
        L.bseg_last_error.argtypes = [vp]
        L.bseg_last_error.restype = C.c_char_p
        L.bseg_version.restype = C.c_char_p
        L.bseg_set_points.argtypes = [vp, vp, i64, vp, vp, vp]
        L.bseg_set_points_device.argtypes = [vp, vp, i64, vp, vp]
        L.bseg_knn_normals.argtypes = [vp, C.POINTER(Params), vp, vp, vp]
        L.bseg_override_neigh_normals.argtypes = [vp, C.POINTER(Params), vp, vp]
        L.bseg_grow_planes.argtypes = [vp, C.POINTER(Params), vp, vp, vp]
        L.bseg_get_planes.argtypes = [vp, vp, vp, vp, vp, vp]
        L.bseg_paint.argtypes = [vp, vp, C.c_int32, vp, vp]
        L.bseg_raster_size.argtypes = [vp, C.POINTER(Params), vp, vp]
        L.bseg_raster.argtypes = [vp, C.POINTER(Params), vp, vp, vp, vp, vp]
        L.bseg_label_raster.argtypes = [vp, C.POINTER(Params), vp, vp, vp]
        L.bseg_raster_device.argtypes = [vp, C.POINTER(Params), vp, vp, vp, vp]
        L.bseg_run_device.argtypes = [vp, C.POINTER(Params), C.c_int]
        L.bseg_segment_host.argtypes = [vp, C.POINTER(Params), vp, i64, vp, vp, vp, vp, vp, vp, vp]
        L.bseg_get_timings.argtypes = [vp, C.POINTER(Timings)]
        L.bseg_reset_counters.argtypes = [vp]
        L.bseg_stream.argtypes = [vp]
        L.bseg_stream.restype = vp
        L.bseg_point_count.argtypes = [vp]
        L.bseg_point_count.restype = i64
        L.bseg_plane_count.argtypes = [vp]
        L.bseg_plane_count.restype = C.c_int32
        L.bseg_set_owned.argtypes = [vp, i64]
        L.bseg_set_origin.argtypes = [vp, vp]
        L.bseg_device_results.argtypes = [vp, vp, vp, vp]
        L.bseg_halo_check.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp]
        L.bseg_count_channel.argtypes = [vp, i64, C.c_double, vp]
        L.bseg_contour_mask.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
        L.bseg_find_contours.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, i64, vp, i64, vp, vp]
        L.bseg_contour_measure.argtypes = [vp, i64, vp, vp]
        L.bseg_draw_contour.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, i64, vp]
        L.bseg_plane_classes.argtypes = [vp, C.c_double, C.c_double, C.c_double, vp, vp, vp]
        L.bseg_png_encode.argtypes = [vp, i32, i32, i32, i32, vp, i64, vp]
        L.bseg_png_write.argtypes = [C.c_char_p, vp, i32, i32, i32, i32]
        L.bseg_png_write_async.argtypes = [C.c_char_p, vp, i32, i32, i32, i32]
        L.bseg_png_wait.argtypes = []
        L.bseg_set_grow_offset.argtypes = [vp, vp]
        L.bseg_knn_device_results.argtypes = [vp, C.POINTER(Params), vp, vp]
        L.bseg_import_neigh_normals_device.argtypes = [vp, C.POINTER(Params), vp, vp]
        L.bseg_debug_sort_pairs.argtypes = [vp, vp, vp, i64, C.c_int]
        L.bseg_debug_exclusive_scan.argtypes = [vp, vp, i64]
        _LIB = L
    return _LIB


def default_params(**kw) -> Params:
    p = Params()
    lib().bseg_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k}")
        setattr(p, k, v)
    return p


def find_contours(mask: np.ndarray, simple: bool = True):
    """bseg_find_contours: list of int32 [k][2] arrays, as cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE/NONE)."""
    mask = np.ascontiguousarray(mask, np.uint8)
    H, W = mask.shape
    nc, npt = C.c_int64(0), C.c_int64(0)
    rc = lib().bseg_find_contours(mask.ctypes.data, W, H, int(simple), None, 0, None, 0, C.addressof(nc), C.addressof(npt))
    if rc != 0:
        raise BsegError(rc, last_error(None))
    pts = np.empty((max(npt.value, 1), 2), np.int32)
    off = np.empty(nc.value + 1, np.int64)
    rc = lib().bseg_find_contours(mask.ctypes.data, W, H, int(simple), pts.ctypes.data, npt.value, off.ctypes.data, nc.value,
                                  C.addressof(nc), C.addressof(npt))
    if rc != 0:
        raise BsegError(rc, last_error(None))
    return [pts[off[k]: off[k + 1]].copy() for k in range(nc.value)]


def contour_measure(pts: np.ndarray):
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    a, l = C.c_double(0.0), C.c_double(0.0)
    rc = lib().bseg_contour_measure(pts.ctypes.data, len(pts), C.addressof(a), C.addressof(l))
    if rc != 0:
        raise BsegError(rc, last_error(None))
    return a.value, l.value


def draw_contour(image: np.ndarray, pts: np.ndarray, color=(0, 255, 255)):
    """bseg_draw_contour in place on a C-contiguous uint8 [H][W][comp] image."""
    assert image.dtype == np.uint8 and image.flags.c_contiguous
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    col = np.asarray(color, np.uint8)
    H, W = image.shape[:2]
    comp = 1 if image.ndim == 2 else image.shape[2]
    rc = lib().bseg_draw_contour(image.ctypes.data, W, H, comp, pts.ctypes.data, len(pts), col.ctypes.data)
    if rc != 0:
        raise BsegError(rc, last_error(None))


def count_channel(values: np.ndarray, bias: float) -> float:
    """bseg_count_channel in place on a C-contiguous float64 array; returns the maximum."""
    assert values.dtype == np.float64 and values.flags.c_contiguous
    m = C.c_double(0.0)
    rc = lib().bseg_count_channel(values.ctypes.data, values.size, float(bias), C.addressof(m))
    if rc != 0:
        raise BsegError(rc, lib().bseg_last_error(None).decode())
    return float(m.value)


def png_encode(img: np.ndarray) -> bytes:
    """bseg_png_encode of a uint8 [H][W][comp] (or [H][W]) image: the bytes stbi_write_png would write."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    comp = 1 if img.ndim == 2 else img.shape[2]
    n = C.c_int64(0)
    rc = lib().bseg_png_encode(img.ctypes.data, w, h, comp, 0, None, 0, C.addressof(n))
    if rc != 0:
        raise BsegError(rc, lib().bseg_last_error(None).decode())
    out = np.empty(n.value, np.uint8)
    rc = lib().bseg_png_encode(img.ctypes.data, w, h, comp, 0, out.ctypes.data, out.size, C.addressof(n))
    if rc != 0:
        raise BsegError(rc, lib().bseg_last_error(None).decode())
    return out.tobytes()


def _ptr(a):
    return None if a is None else a.ctypes.data


RUN_KNN, RUN_GROW, RUN_RASTER, RUN_ALL = 1, 2, 4, 7
CLASS_NONE, CLASS_ROOF, CLASS_FACADE, CLASS_GROUND, CLASS_OTHER = range(5)  # bseg_plane_classes


class Context:
    """One libbseg context = one CUDA device + one stream.  Mirrors the call order of the reference's
    main (TMC3.cpp:202-229): set_points -> knn_normals -> grow_planes -> paint / raster."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().bseg_create(C.byref(self._h), device)
        if rc != 0:
            raise BsegError(rc, lib().bseg_last_error(None).decode())
        self.n = 0

    def close(self):
        if self._h:
            lib().bseg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise BsegError(rc, lib().bseg_last_error(self._h).decode())

    # -- a3 ---------------------------------------------------------------------------------------
    def set_points(self, xyz, want_shifted=True):
        xyz = np.ascontiguousarray(xyz, np.int32)
        assert xyz.ndim == 2 and xyz.shape[1] == 3
        self.n = len(xyz)
        mn = np.zeros(3, np.int32)
        mx = np.zeros(3, np.int32)
        out = np.empty_like(xyz) if want_shifted else None
        self._ck(lib().bseg_set_points(self._h, _ptr(xyz), self.n, _ptr(mn), _ptr(mx), _ptr(out)))
        return mn, mx, out

    def set_points_device(self, dev_ptr, n):
        self.n = int(n)
        mn = np.zeros(3, np.int32)
        mx = np.zeros(3, np.int32)
        self._ck(lib().bseg_set_points_device(self._h, dev_ptr, self.n, _ptr(mn), _ptr(mx)))
        return mn, mx

    def set_owned(self, n_owned):
        self._ck(lib().bseg_set_owned(self._h, int(n_owned)))

    def n_planes(self):
        return int(lib().bseg_plane_count(self._h))

    def set_origin(self, origin):
        """Shift the next clouds by `origin` (3 ints, the tile's minimum) instead of their own minimum; None resets."""
        if origin is None:
            self._ck(lib().bseg_set_origin(self._h, None))
        else:
            o = np.ascontiguousarray(origin, np.int32)
            self._ck(lib().bseg_set_origin(self._h, o.ctypes.data))

    def device_results(self):
        """Device addresses (ints) of label[n], planeIdx[n] (original order) and the shifted cloud [n][3]."""
        a, b, c_ = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(lib().bseg_device_results(self._h, C.addressof(a), C.addressof(b), C.addressof(c_)))
        return a.value, b.value, c_.value

    def halo_check(self, x_lo, x_hi, halo):
        n = C.c_int64(0)
        self._ck(lib().bseg_halo_check(self._h, int(x_lo), int(x_hi), int(halo), C.addressof(n)))
        return int(n.value)

    # -- a4 + a5 ----------------------------------------------------------------------------------
    def knn_normals(self, p: Params, want_neigh=True, want_normals=True, want_curvature=False):
        n = self.n
        neigh = np.empty((n, p.K), np.int32) if want_neigh else None
        nrm = np.empty((n, 3), np.float64) if want_normals else None
        curv = np.empty(n, np.float64) if want_curvature else None
        self._ck(lib().bseg_knn_normals(self._h, C.byref(p), _ptr(neigh), _ptr(nrm), _ptr(curv)))
        return neigh, nrm, curv

    def override(self, p: Params, neigh=None, normals=None):
        if neigh is not None:
            neigh = np.ascontiguousarray(neigh, np.int32)
            assert neigh.shape == (self.n, p.K)
        if normals is not None:
            normals = np.ascontiguousarray(normals, np.float64)
            assert normals.shape == (self.n, 3)
        self._ck(lib().bseg_override_neigh_normals(self._h, C.byref(p), _ptr(neigh), _ptr(normals)))

    def knn_device_results(self, p: Params):
        """Device addresses of the rows [n][K] (int32, original indices) and normals [n][3] (float64) of the last kNN stage."""
        a, b = C.c_void_p(), C.c_void_p()
        self._ck(lib().bseg_knn_device_results(self._h, C.byref(p), C.addressof(a), C.addressof(b)))
        return a.value, b.value

    def import_neigh_normals_device(self, p: Params, d_neigh, d_normals):
        """Rows / normals computed elsewhere (device pointers, original indexing) become the grower's input; no kNN runs."""
        self._ck(lib().bseg_import_neigh_normals_device(self._h, C.byref(p), int(d_neigh), int(d_normals)))

    def set_grow_offset(self, offset):
        if offset is None:
            self._ck(lib().bseg_set_grow_offset(self._h, None))
        else:
            o = np.ascontiguousarray(offset, np.int32)
            self._ck(lib().bseg_set_grow_offset(self._h, o.ctypes.data))

    # -- a7-a9 ------------------------------------------------------------------------------------
    def grow_planes(self, p: Params, want_arrays=True):
        n = self.n
        pidx = np.empty(n, np.int32) if want_arrays else None
        label = np.empty(n, np.int32) if want_arrays else None
        npl = C.c_int32(0)
        self._ck(lib().bseg_grow_planes(self._h, C.byref(p), _ptr(pidx), _ptr(label), C.addressof(npl)))
        return pidx, label, int(npl.value)

    def get_planes(self, n_planes):
        P = int(n_planes)
        seeds = np.empty(P, np.int32)
        normals = np.empty((P, 3), np.float64)
        centers = np.empty((P, 3), np.int32)
        off = np.zeros(P + 1, np.int64)
        self._ck(lib().bseg_get_planes(self._h, _ptr(seeds), _ptr(normals), _ptr(centers), _ptr(off), None))
        idx = np.empty(int(off[P]), np.int32)
        self._ck(lib().bseg_get_planes(self._h, _ptr(seeds), _ptr(normals), _ptr(centers), _ptr(off), _ptr(idx)))
        return seeds, normals, centers, off, idx

    # -- a10 --------------------------------------------------------------------------------------
    def paint(self, plane_rgb, plane_ids=None):
        """set_plane_color: all planes in id order (plane_ids None), or the listed planes in the listed order."""
        plane_rgb = np.ascontiguousarray(plane_rgb, np.uint16).reshape(-1, 3)
        colors = np.empty((self.n, 3), np.uint16)
        ids = None if plane_ids is None else np.ascontiguousarray(plane_ids, np.int32)
        self._ck(lib().bseg_paint(self._h, None if ids is None else _ptr(ids), len(plane_rgb), _ptr(plane_rgb), _ptr(colors)))
        return colors

    def plane_classes(self, ground_z, facade_max_nz=0.3, roof_min_nz=0.7, want_points=True):
        """DetectedPlane equations (my_function.h:41-46) and roof / facade / ground classes of the planes of the last
        grow: (equations [P][4] float64, plane_class [P] uint8, point_class [N] uint8 or None)."""
        P = self.n_planes()
        eq = np.empty((P, 4), np.float64)
        pc = np.empty(P, np.uint8)
        pt = np.empty(self.n, np.uint8) if want_points else None
        self._ck(lib().bseg_plane_classes(self._h, float(facade_max_nz), float(roof_min_nz), float(ground_z), _ptr(eq), _ptr(pc),
                                          _ptr(pt)))
        return eq, pc, pt

    def contour_mask(self, pixels: np.ndarray, channel=1, thresh=10, iterations=2):
        """bseg_contour_mask: threshold + morphological close on the device; pixels uint8 [H][W][comp] (host)."""
        pixels = np.ascontiguousarray(pixels, np.uint8)
        H, W = pixels.shape[:2]
        comp = 1 if pixels.ndim == 2 else pixels.shape[2]
        out = np.empty((H, W), np.uint8)
        self._ck(lib().bseg_contour_mask(self._h, _ptr(pixels), W, H, comp, channel, thresh, iterations, _ptr(out)))
        return out

    # -- a13-a15 ----------------------------------------------------------------------------------
    def raster_size(self, p: Params):
        W = C.c_int32(0)
        H = C.c_int32(0)
        self._ck(lib().bseg_raster_size(self._h, C.byref(p), C.addressof(W), C.addressof(H)))
        return int(W.value), int(H.value)

    def raster(self, p: Params, want_image=True, want_png=True):
        W, H = self.raster_size(p)
        img = np.empty((H, W, 3), np.float64) if want_image else None
        a = np.empty((H, W, 3), np.uint8) if want_png else None
        b = np.empty((H, W, 3), np.uint8) if want_png else None
        c = np.empty((H, W, 3), np.uint8) if want_png else None
        th = C.c_double(0.0)
        self._ck(lib().bseg_raster(self._h, C.byref(p), _ptr(img), _ptr(a), _ptr(b), _ptr(c), C.addressof(th)))
        return img, a, b, c, float(th.value)

    def raster_device(self, p: Params, ground_th=None):
        """bseg_raster_device: (device pointer of the W*H*3 doubles, W, H); ground_th overrides groundTH."""
        th = C.c_double(0.0 if ground_th is None else float(ground_th))
        d_img = C.c_void_p()
        W = C.c_int32(0)
        H = C.c_int32(0)
        self._ck(lib().bseg_raster_device(self._h, C.byref(p), None if ground_th is None else C.addressof(th),
                                          C.addressof(d_img), C.addressof(W), C.addressof(H)))
        return int(d_img.value or 0), int(W.value), int(H.value)

    def label_raster(self, p: Params, plane_rgb=None):
        """Plane label of the highest point per pixel + its colour image (bseg_label_raster)."""
        W, H = self.raster_size(p)
        lab = np.empty((H, W), np.int32)
        rgb = np.empty((H, W, 3), np.uint8)
        if plane_rgb is not None:
            plane_rgb = np.ascontiguousarray(plane_rgb, np.uint16)
        self._ck(lib().bseg_label_raster(self._h, C.byref(p), _ptr(plane_rgb), _ptr(lab), _ptr(rgb)))
        return lab, rgb

    # -- whole path -------------------------------------------------------------------------------
    def run_device(self, p: Params, stages=RUN_ALL):
        self._ck(lib().bseg_run_device(self._h, C.byref(p), stages))

    def segment_host(self, p: Params, xyz, shifted_out, label_out, png_a=None, png_b=None):
        """bseg_segment_host on caller-owned (ideally pinned) numpy buffers; returns (n_planes, W, H)."""
        npl = C.c_int32(0)
        W = C.c_int32(0)
        H = C.c_int32(0)
        self.n = len(xyz)
        self._ck(lib().bseg_segment_host(self._h, C.byref(p), _ptr(xyz), len(xyz), _ptr(shifted_out), _ptr(label_out),
                                         C.addressof(npl), _ptr(png_a), _ptr(png_b), C.addressof(W), C.addressof(H)))
        return int(npl.value), int(W.value), int(H.value)

    def timings(self) -> dict:
        t = Timings()
        self._ck(lib().bseg_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def reset_counters(self):
        self._ck(lib().bseg_reset_counters(self._h))

    @property
    def stream(self):
        return lib().bseg_stream(self._h)

    # -- self-test hooks --------------------------------------------------------------------------
    def debug_sort_pairs(self, keys, vals, key_bits):
        keys = np.ascontiguousarray(keys, np.uint64).copy()
        vals = np.ascontiguousarray(vals, np.uint32).copy()
        self._ck(lib().bseg_debug_sort_pairs(self._h, _ptr(keys), _ptr(vals), len(keys), key_bits))
        return keys, vals

    def debug_exclusive_scan(self, data):
        data = np.ascontiguousarray(data, np.uint32).copy()
        self._ck(lib().bseg_debug_exclusive_scan(self._h, _ptr(data), len(data)))
        return data
