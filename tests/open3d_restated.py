"""An INDEPENDENT numpy transcription of the Open3D 0.19.0 arithmetic the reference calls at
/root/reference/tmc3/my_function.h:63-64 (EstimateNormals(Hybrid) + OrientNormalsToAlignWithDirection):

    utility::ComputeCovariance          cpp/open3d/utility/Eigen.cpp  (cumulants in list order, /= n, E[ab] - E[a]E[b])
    utility::FastEigen3x3               cpp/open3d/geometry/EstimateNormals.cpp  (Eberly, "A Robust Eigensolver for 3x3
    ComputeEigenvector0 / 1              Symmetric Matrices", geometrictools.com)
    PointCloud::EstimateNormals         zero normal -> (0,0,1)
    OrientNormalsToAlignWithDirection   n . d < 0 -> -n

TEST INFRASTRUCTURE.  It deliberately shares NOTHING with the product or the C oracle: it does not include
buildingsegment_b200/csrc/bseg_arith.h, it is written array-at-a-time (every branch of the scalar algorithm becomes a
mask), and the two transcendental calls are parameters (`acos`, `cos`): numpy's by default, or the oracle's exported
polynomial kernels when a test wants to compare bit patterns rather than values.

Open3D is not installable here (no network) and the reference has no golden vectors, so this remains a transcription
from the published sources -- but a second, separately written one: a slip in either shows up as a disagreement.
Every elementary operation below is a single IEEE-754 double operation in the order the C++ evaluates it.
"""
from __future__ import annotations

import numpy as np

TWO_THIRDS_PI = 2.09439510239319549


def covariance_from_points(pts: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """utility::ComputeCovariance for M neighbour lists.  pts: float64 [M][L][3] (entries beyond counts[m] ignored),
    counts: [M] (>= 1).  Returns the symmetric matrices as [M][3][3]."""
    M, L, _ = pts.shape
    cum = np.zeros((M, 9), np.float64)
    for k in range(L):  # list order: cumulants(i) += ... one neighbour at a time
        live = k < counts
        x = np.where(live, pts[:, k, 0], 0.0)
        y = np.where(live, pts[:, k, 1], 0.0)
        z = np.where(live, pts[:, k, 2], 0.0)
        cum[:, 0] += x
        cum[:, 1] += y
        cum[:, 2] += z
        cum[:, 3] += x * x
        cum[:, 4] += x * y
        cum[:, 5] += x * z
        cum[:, 6] += y * y
        cum[:, 7] += y * z
        cum[:, 8] += z * z
    cum = cum / counts.astype(np.float64)[:, None]
    C = np.empty((M, 3, 3), np.float64)
    C[:, 0, 0] = cum[:, 3] - cum[:, 0] * cum[:, 0]
    C[:, 1, 1] = cum[:, 6] - cum[:, 1] * cum[:, 1]
    C[:, 2, 2] = cum[:, 8] - cum[:, 2] * cum[:, 2]
    C[:, 0, 1] = cum[:, 4] - cum[:, 0] * cum[:, 1]
    C[:, 1, 0] = C[:, 0, 1]
    C[:, 0, 2] = cum[:, 5] - cum[:, 0] * cum[:, 2]
    C[:, 2, 0] = C[:, 0, 2]
    C[:, 1, 2] = cum[:, 7] - cum[:, 1] * cum[:, 2]
    C[:, 2, 1] = C[:, 1, 2]
    return C


def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                     a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def _dot(a, b):
    return a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]


def eigenvector0(A: np.ndarray, ev: np.ndarray) -> np.ndarray:
    """ComputeEigenvector0: rows of A - ev*I, the pairwise cross product of largest squared norm, normalised."""
    r0 = np.stack([A[:, 0, 0] - ev, A[:, 0, 1], A[:, 0, 2]], axis=1)
    r1 = np.stack([A[:, 0, 1], A[:, 1, 1] - ev, A[:, 1, 2]], axis=1)
    r2 = np.stack([A[:, 0, 2], A[:, 1, 2], A[:, 2, 2] - ev], axis=1)
    r0xr1, r0xr2, r1xr2 = _cross(r0, r1), _cross(r0, r2), _cross(r1, r2)
    d0, d1, d2 = _dot(r0xr1, r0xr1), _dot(r0xr2, r0xr2), _dot(r1xr2, r1xr2)
    dmax = d0.copy()
    imax = np.zeros(len(ev), np.int64)
    m = d1 > dmax
    dmax = np.where(m, d1, dmax)
    imax = np.where(m, 1, imax)
    m = d2 > dmax
    imax = np.where(m, 2, imax)
    with np.errstate(all="ignore"):
        v0 = r0xr1 / np.sqrt(d0)[:, None]
        v1 = r0xr2 / np.sqrt(d1)[:, None]
        v2 = r1xr2 / np.sqrt(d2)[:, None]
    return np.where((imax == 0)[:, None], v0, np.where((imax == 1)[:, None], v1, v2))


def eigenvector1(A: np.ndarray, e0: np.ndarray, ev1: np.ndarray) -> np.ndarray:
    """ComputeEigenvector1: an eigenvector for ev1 in the plane orthogonal to e0."""
    with np.errstate(all="ignore"):
        big_x = np.abs(e0[:, 0]) > np.abs(e0[:, 1])
        inv_a = 1.0 / np.sqrt(e0[:, 0] * e0[:, 0] + e0[:, 2] * e0[:, 2])
        inv_b = 1.0 / np.sqrt(e0[:, 1] * e0[:, 1] + e0[:, 2] * e0[:, 2])
        zero = np.zeros_like(ev1)
        U = np.where(big_x[:, None], np.stack([-e0[:, 2] * inv_a, zero, e0[:, 0] * inv_a], 1),
                     np.stack([zero, e0[:, 2] * inv_b, -e0[:, 1] * inv_b], 1))
        V = _cross(e0, U)
        AU = np.stack([A[:, 0, 0] * U[:, 0] + A[:, 0, 1] * U[:, 1] + A[:, 0, 2] * U[:, 2],
                       A[:, 0, 1] * U[:, 0] + A[:, 1, 1] * U[:, 1] + A[:, 1, 2] * U[:, 2],
                       A[:, 0, 2] * U[:, 0] + A[:, 1, 2] * U[:, 1] + A[:, 2, 2] * U[:, 2]], 1)
        AV = np.stack([A[:, 0, 0] * V[:, 0] + A[:, 0, 1] * V[:, 1] + A[:, 0, 2] * V[:, 2],
                       A[:, 0, 1] * V[:, 0] + A[:, 1, 1] * V[:, 1] + A[:, 1, 2] * V[:, 2],
                       A[:, 0, 2] * V[:, 0] + A[:, 1, 2] * V[:, 1] + A[:, 2, 2] * V[:, 2]], 1)
        m00 = _dot(U, AU) - ev1
        m01 = _dot(U, AV)
        m11 = _dot(V, AV) - ev1
        a00, a01, a11 = np.abs(m00), np.abs(m01), np.abs(m11)

        # branch A: |m00| >= |m11|
        maxA = np.maximum(a00, a01)
        # A1: |m00| >= |m01|   m01 /= m00; m00 = 1/sqrt(1+m01^2); m01 *= m00
        t = m01 / m00
        c = 1.0 / np.sqrt(1.0 + t * t)
        rA1 = (t * c)[:, None] * U - c[:, None] * V
        # A2: m00 /= m01; m01 = 1/sqrt(1+m00^2); m00 *= m01
        t = m00 / m01
        c = 1.0 / np.sqrt(1.0 + t * t)
        rA2 = c[:, None] * U - (t * c)[:, None] * V
        rA = np.where((maxA > 0)[:, None], np.where((a00 >= a01)[:, None], rA1, rA2), U)

        # branch B: |m00| < |m11|
        maxB = np.maximum(a11, a01)
        # B1: |m11| >= |m01|   m01 /= m11; m11 = 1/sqrt(1+m01^2); m01 *= m11
        t = m01 / m11
        c = 1.0 / np.sqrt(1.0 + t * t)
        rB1 = c[:, None] * U - (t * c)[:, None] * V
        # B2: m11 /= m01; m01 = 1/sqrt(1+m11^2); m11 *= m01
        t = m11 / m01
        c = 1.0 / np.sqrt(1.0 + t * t)
        rB2 = (t * c)[:, None] * U - c[:, None] * V
        rB = np.where((maxB > 0)[:, None], np.where((a11 >= a01)[:, None], rB1, rB2), U)
    return np.where((a00 >= a11)[:, None], rA, rB)


def fast_eigen3x3(C: np.ndarray, acos=np.arccos, cos=np.cos) -> np.ndarray:
    """FastEigen3x3: unit eigenvector of the smallest eigenvalue of every symmetric matrix in C ([M][3][3]);
    the zero vector where the matrix has no positive coefficient."""
    M = len(C)
    out = np.zeros((M, 3), np.float64)
    max_coeff = C.reshape(M, 9).max(axis=1)
    ok = max_coeff != 0
    with np.errstate(all="ignore"):
        A = C / np.where(ok, max_coeff, 1.0)[:, None, None]
        norm = A[:, 0, 1] * A[:, 0, 1] + A[:, 0, 2] * A[:, 0, 2] + A[:, 1, 2] * A[:, 1, 2]
        gen = ok & (norm > 0)

        q = (A[:, 0, 0] + A[:, 1, 1] + A[:, 2, 2]) / 3.0
        b00, b11, b22 = A[:, 0, 0] - q, A[:, 1, 1] - q, A[:, 2, 2] - q
        p = np.sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0)
        c00 = b11 * b22 - A[:, 1, 2] * A[:, 1, 2]
        c01 = A[:, 0, 1] * b22 - A[:, 1, 2] * A[:, 0, 2]
        c02 = A[:, 0, 1] * A[:, 1, 2] - b11 * A[:, 0, 2]
        det = (b00 * c00 - A[:, 0, 1] * c01 + A[:, 0, 2] * c02) / (p * p * p)
        half_det = det * 0.5
        half_det = np.where(half_det < -1.0, -1.0, half_det)  # std::min(std::max(h, -1.0), 1.0)
        half_det = np.where(half_det > 1.0, 1.0, half_det)
        angle = acos(half_det) / 3.0
        beta2 = cos(angle) * 2.0
        beta0 = cos(angle + TWO_THIRDS_PI) * 2.0
        beta1 = -(beta0 + beta2)
        ev0, ev1, ev2 = q + p * beta0, q + p * beta1, q + p * beta2

        # half_det >= 0: start from the largest eigenvalue
        e2 = eigenvector0(A, ev2)
        e1p = eigenvector1(A, e2, ev1)
        pos = np.where(((ev2 < ev0) & (ev2 < ev1))[:, None], e2,
                       np.where(((ev1 < ev0) & (ev1 < ev2))[:, None], e1p, _cross(e1p, e2)))
        # half_det < 0: start from the smallest
        e0 = eigenvector0(A, ev0)
        e1n = eigenvector1(A, e0, ev1)
        neg = np.where(((ev0 < ev1) & (ev0 < ev2))[:, None], e0,
                       np.where(((ev1 < ev0) & (ev1 < ev2))[:, None], e1n, _cross(e0, e1n)))
        general = np.where((half_det >= 0)[:, None], pos, neg)

    # diagonal matrix: the axis of the strictly smallest diagonal entry, else z
    ax = np.zeros((M, 3), np.float64)
    sx = (A[:, 0, 0] < A[:, 1, 1]) & (A[:, 0, 0] < A[:, 2, 2])
    sy = ~sx & (A[:, 1, 1] < A[:, 0, 0]) & (A[:, 1, 1] < A[:, 2, 2])
    ax[sx, 0] = 1.0
    ax[sy, 1] = 1.0
    ax[~sx & ~sy, 2] = 1.0
    out = np.where(gen[:, None], general, np.where(ok[:, None], ax, 0.0))
    return out


def normals_from_covariances(C: np.ndarray, acos=np.arccos, cos=np.cos) -> np.ndarray:
    """EstimateNormals' per-point tail + OrientNormalsToAlignWithDirection((0,0,1))."""
    n = fast_eigen3x3(C, acos, cos)
    nn = n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2]
    zero = nn == 0.0  # Eigen norm() == 0; NaN stays NaN, as upstream
    n = np.where(zero[:, None], np.array([0.0, 0.0, 1.0]), n)
    flip = n[:, 2] < 0.0
    return np.where(flip[:, None], n * -1.0, n)


def estimate_normals(xyz: np.ndarray, knn_idx: np.ndarray, knn_d2: np.ndarray, radius=100.0, max_nn=50,
                     acos=np.arccos, cos=np.cos) -> np.ndarray:
    """my_function.h:63-64 for a whole cloud, given the ordered kNN rows: hybrid set = the first min(max_nn,
    #{d^2 < r^2}) entries; fewer than 3 -> identity covariance."""
    n = len(xyz)
    lim = min(max_nn, knn_idx.shape[1])
    inside = (knn_idx[:, :lim] >= 0) & (knn_d2[:, :lim].astype(np.float64) < radius * radius)
    # prefix length (the d^2 are ascending, but count the prefix explicitly)
    cnt = np.where(inside.all(axis=1), lim, np.argmin(inside, axis=1)).astype(np.int64)
    pts = xyz[np.clip(knn_idx[:, :lim], 0, n - 1)].astype(np.float64)
    C = covariance_from_points(pts, np.maximum(cnt, 1))
    few = cnt < 3
    C[few] = np.eye(3)
    return normals_from_covariances(C, acos, cos)
