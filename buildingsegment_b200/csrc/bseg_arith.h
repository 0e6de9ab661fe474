/*
 * bseg_arith.h -- shared IEEE-754-only fp64 arithmetic for the segmentation hot path.
 *
 * Compiled three ways: CUDA device code (nvcc -fmad=false), the C++ host, and the C oracle
 * (gcc -ffp-contract=off).  Every routine is built from +, -, *, /, sqrt and integer bit
 * moves only, evaluated in one fixed order, so device and host produce the same bits.
 * No libm transcendental is called: glibc and CUDA round acos/cos/log differently, and the
 * plane grower's tests (n.n_i >= 0.88, the sign-of-n_z flip) are discontinuous in the normal.
 *
 * What it restates (the arithmetic lives in Open3D 0.19.0, which the reference links but does
 * not vendor -- call sites /root/reference/tmc3/my_function.h:63-64):
 *   - utility::ComputeCovariance      -> bseg_covariance_from_sums
 *   - utility::FastEigen3x3 (+ ComputeEigenvector0/1, Eberly's robust symmetric 3x3)
 *                                     -> bseg_fast_eigen3x3
 *   - PointCloud::EstimateNormals zero-normal rule and
 *     OrientNormalsToAlignWithDirection((0,0,1))   -> bseg_normal_from_cov
 * and the raster's std::log (TMC3.cpp:161) -> bseg_log.
 * acos/cos/log follow the classic fdlibm minimax kernels (|err| < 1 ulp).
 */
#ifndef BSEG_ARITH_H
#define BSEG_ARITH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define BSEG_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define BSEG_HD static inline
#endif

BSEG_HD uint64_t bseg_d2u(double x)
{
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, sizeof(u));
  return u;
#endif
}

BSEG_HD double bseg_u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, sizeof(x));
  return x;
#endif
}

BSEG_HD double bseg_sqrt(double x)
{
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return sqrt(x);
#endif
}

BSEG_HD double bseg_fabs(double x)
{
  return bseg_u2d(bseg_d2u(x) & 0x7fffffffffffffffULL);
}

/* ---- acos on [-1, 1] -------------------------------------------------------------------- */
BSEG_HD double bseg_acos(double x)
{
  const double pi = 3.14159265358979311600e+00;
  const double pio2_hi = 1.57079632679489655800e+00;
  const double pio2_lo = 6.12323399573676603587e-17;
  const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
               pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
               pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
  const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
               qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
  double ax = bseg_fabs(x);
  if (ax >= 1.0) {
    if (x == 1.0)
      return 0.0;
    if (x == -1.0)
      return pi;
    return (x - x) / (x - x); /* NaN, as libm */
  }
  if (ax < 0.5) {
    if (ax < 6.938893903907228e-18) /* 2^-57 */
      return pio2_hi + pio2_lo;
    double z = x * x;
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    double r = p / q;
    return pio2_hi - (x - (pio2_lo - r * x));
  }
  if (x < 0.0) {
    double z = (1.0 + x) * 0.5;
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    double s = bseg_sqrt(z);
    double r = p / q;
    double w = r * s - pio2_lo;
    return pi - 2.0 * (s + w);
  }
  {
    double z = (1.0 - x) * 0.5;
    double s = bseg_sqrt(z);
    double df = bseg_u2d(bseg_d2u(s) & 0xffffffff00000000ULL);
    double c = (z - df * df) / (s + df);
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    double r = p / q;
    double w = r * s + c;
    return 2.0 * (df + w);
  }
}

/* ---- cos on [0, 4] (the eigen solver needs [0, pi]) --------------------------------------- */
BSEG_HD double bseg_kcos(double x, double y)
{
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double ax = bseg_fabs(x);
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  if (ax < 0.30000000000000004)
    return 1.0 - (0.5 * z - (z * r - x * y));
  double qx;
  if (ax > 0.78125)
    qx = 0.28125;
  else
    qx = bseg_u2d(((bseg_d2u(ax) >> 32) - 0x00200000ULL) << 32); /* ~ |x|/4, low word 0 */
  double hz = 0.5 * z - qx;
  double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}

BSEG_HD double bseg_ksin(double x, double y)
{
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = x * x;
  double v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

BSEG_HD double bseg_cos(double x)
{
  /* pi/2 split: pio2_1 carries 33 bits, pio2_1t the next 53 -- exact n*pio2_1 for n <= 3 */
  const double invpio2 = 6.36619772367581382433e-01;
  const double pio2_1 = 1.57079632673412561417e+00;
  const double pio2_1t = 6.07710050650619224932e-11;
  if (x < 0.7853981633974483)
    return bseg_kcos(x, 0.0);
  int n = (int)(x * invpio2 + 0.5);
  double fn = (double)n;
  double r = x - fn * pio2_1;
  double w = fn * pio2_1t;
  double y0 = r - w;
  double y1 = (r - y0) - w;
  switch (n & 3) {
  case 0: return bseg_kcos(y0, y1);
  case 1: return -bseg_ksin(y0, y1);
  case 2: return -bseg_kcos(y0, y1);
  default: return bseg_ksin(y0, y1);
  }
}

/* ---- natural log for finite x >= 1 (the raster only ever takes log(count + 1)) ------------- */
BSEG_HD double bseg_log(double x)
{
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  uint64_t u = bseg_d2u(x);
  int k = (int)((u >> 52) & 0x7ff) - 1023;
  uint32_t hx = (uint32_t)((u >> 32) & 0x000fffff); /* top 20 mantissa bits */
  uint32_t i = (hx + 0x95f64) & 0x100000;           /* mantissa above sqrt(2)? */
  /* normalise x to [sqrt(2)/2, sqrt(2)) */
  u = (u & 0x000fffffffffffffULL) | ((uint64_t)(i ^ 0x3ff00000) << 32);
  k += (int)(i >> 20);
  double f = bseg_u2d(u) - 1.0;
  double dk = (double)k;
  if (f == 0.0)
    return dk * ln2_hi + dk * ln2_lo;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  int32_t ii = (int32_t)hx - 0x6147a;
  int32_t jj = 0x6b851 - (int32_t)hx;
  if ((ii | jj) > 0) {
    double hfsq = 0.5 * f * f;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  }
  return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

/* ---- covariance from integer-exact moment sums -------------------------------------------
 * sums = { Sx, Sy, Sz, Sxx, Sxy, Sxz, Syy, Syz, Szz } over the n neighbours; every sum is an
 * exact integer below 2^53, so the summation order of Open3D's loop is immaterial and the
 * first rounding happens at the division by n, exactly as in utility::ComputeCovariance.
 * cov = { c00, c01, c02, c11, c12, c22 }.
 */
BSEG_HD void bseg_covariance_from_sums(const double sums[9], int n, double cov[6])
{
  double dn = (double)n;
  double m0 = sums[0] / dn, m1 = sums[1] / dn, m2 = sums[2] / dn;
  double e00 = sums[3] / dn, e01 = sums[4] / dn, e02 = sums[5] / dn;
  double e11 = sums[6] / dn, e12 = sums[7] / dn, e22 = sums[8] / dn;
  cov[0] = e00 - m0 * m0;
  cov[1] = e01 - m0 * m1;
  cov[2] = e02 - m0 * m2;
  cov[3] = e11 - m1 * m1;
  cov[4] = e12 - m1 * m2;
  cov[5] = e22 - m2 * m2;
}

/* unit eigenvector of (A - eval*I): largest cross product of two rows (ComputeEigenvector0) */
BSEG_HD void bseg_evec0(double a00, double a01, double a02, double a11, double a12, double a22,
                        double ev, double out[3])
{
  double r0x = a00 - ev, r0y = a01, r0z = a02;
  double r1x = a01, r1y = a11 - ev, r1z = a12;
  double r2x = a02, r2y = a12, r2z = a22 - ev;
  double c01x = r0y * r1z - r0z * r1y, c01y = r0z * r1x - r0x * r1z, c01z = r0x * r1y - r0y * r1x;
  double c02x = r0y * r2z - r0z * r2y, c02y = r0z * r2x - r0x * r2z, c02z = r0x * r2y - r0y * r2x;
  double c12x = r1y * r2z - r1z * r2y, c12y = r1z * r2x - r1x * r2z, c12z = r1x * r2y - r1y * r2x;
  double d0 = c01x * c01x + c01y * c01y + c01z * c01z;
  double d1 = c02x * c02x + c02y * c02y + c02z * c02z;
  double d2 = c12x * c12x + c12y * c12y + c12z * c12z;
  double dmax = d0;
  int imax = 0;
  if (d1 > dmax) {
    dmax = d1;
    imax = 1;
  }
  if (d2 > dmax)
    imax = 2;
  if (imax == 0) {
    double s = bseg_sqrt(d0);
    out[0] = c01x / s; out[1] = c01y / s; out[2] = c01z / s;
  } else if (imax == 1) {
    double s = bseg_sqrt(d1);
    out[0] = c02x / s; out[1] = c02y / s; out[2] = c02z / s;
  } else {
    double s = bseg_sqrt(d2);
    out[0] = c12x / s; out[1] = c12y / s; out[2] = c12z / s;
  }
}

/* second eigenvector in the plane orthogonal to e0 (ComputeEigenvector1) */
BSEG_HD void bseg_evec1(double a00, double a01, double a02, double a11, double a12, double a22,
                        const double e0[3], double ev1, double out[3])
{
  double U0, U1, U2;
  if (bseg_fabs(e0[0]) > bseg_fabs(e0[1])) {
    double inv = 1.0 / bseg_sqrt(e0[0] * e0[0] + e0[2] * e0[2]);
    U0 = -e0[2] * inv; U1 = 0.0; U2 = e0[0] * inv;
  } else {
    double inv = 1.0 / bseg_sqrt(e0[1] * e0[1] + e0[2] * e0[2]);
    U0 = 0.0; U1 = e0[2] * inv; U2 = -e0[1] * inv;
  }
  double V0 = e0[1] * U2 - e0[2] * U1;
  double V1 = e0[2] * U0 - e0[0] * U2;
  double V2 = e0[0] * U1 - e0[1] * U0;
  double AU0 = a00 * U0 + a01 * U1 + a02 * U2;
  double AU1 = a01 * U0 + a11 * U1 + a12 * U2;
  double AU2 = a02 * U0 + a12 * U1 + a22 * U2;
  double AV0 = a00 * V0 + a01 * V1 + a02 * V2;
  double AV1 = a01 * V0 + a11 * V1 + a12 * V2;
  double AV2 = a02 * V0 + a12 * V1 + a22 * V2;
  double m00 = U0 * AU0 + U1 * AU1 + U2 * AU2 - ev1;
  double m01 = U0 * AV0 + U1 * AV1 + U2 * AV2;
  double m11 = V0 * AV0 + V1 * AV1 + V2 * AV2 - ev1;
  double am00 = bseg_fabs(m00), am01 = bseg_fabs(m01), am11 = bseg_fabs(m11);
  if (am00 >= am11) {
    double mx = am00 > am01 ? am00 : am01;
    if (mx > 0.0) {
      if (am00 >= am01) {
        m01 /= m00;
        m00 = 1.0 / bseg_sqrt(1.0 + m01 * m01);
        m01 *= m00;
      } else {
        m00 /= m01;
        m01 = 1.0 / bseg_sqrt(1.0 + m00 * m00);
        m00 *= m01;
      }
      out[0] = m01 * U0 - m00 * V0; out[1] = m01 * U1 - m00 * V1; out[2] = m01 * U2 - m00 * V2;
    } else {
      out[0] = U0; out[1] = U1; out[2] = U2;
    }
  } else {
    double mx = am11 > am01 ? am11 : am01;
    if (mx > 0.0) {
      if (am11 >= am01) {
        m01 /= m11;
        m11 = 1.0 / bseg_sqrt(1.0 + m01 * m01);
        m01 *= m11;
      } else {
        m11 /= m01;
        m01 = 1.0 / bseg_sqrt(1.0 + m11 * m11);
        m11 *= m01;
      }
      out[0] = m11 * U0 - m01 * V0; out[1] = m11 * U1 - m01 * V1; out[2] = m11 * U2 - m01 * V2;
    } else {
      out[0] = U0; out[1] = U1; out[2] = U2;
    }
  }
}

/* eigenvector of the smallest eigenvalue of a symmetric PSD 3x3 (FastEigen3x3);
 * also returns the three eigenvalues of cov in ascending order when evals != NULL. */
BSEG_HD void bseg_fast_eigen3x3(const double cov[6], double out[3], double evals[3])
{
  double a00 = cov[0], a01 = cov[1], a02 = cov[2], a11 = cov[3], a12 = cov[4], a22 = cov[5];
  double mc = a00;
  if (a01 > mc) mc = a01;
  if (a02 > mc) mc = a02;
  if (a11 > mc) mc = a11;
  if (a12 > mc) mc = a12;
  if (a22 > mc) mc = a22;
  if (evals) { evals[0] = 0.0; evals[1] = 0.0; evals[2] = 0.0; }
  if (mc == 0.0) {
    out[0] = 0.0; out[1] = 0.0; out[2] = 0.0;
    return;
  }
  a00 /= mc; a01 /= mc; a02 /= mc; a11 /= mc; a12 /= mc; a22 /= mc;
  double norm = a01 * a01 + a02 * a02 + a12 * a12;
  if (norm > 0.0) {
    double q = (a00 + a11 + a22) / 3.0;
    double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    double p = bseg_sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2.0) / 6.0);
    double c00 = b11 * b22 - a12 * a12;
    double c01 = a01 * b22 - a12 * a02;
    double c02 = a01 * a12 - b11 * a02;
    double det = (b00 * c00 - a01 * c01 + a02 * c02) / (p * p * p);
    double half_det = det * 0.5;
    if (half_det < -1.0) half_det = -1.0; /* std::min(std::max(h,-1),1); NaN passes through */
    if (half_det > 1.0) half_det = 1.0;
    double angle = bseg_acos(half_det) / 3.0;
    const double two_thirds_pi = 2.09439510239319549;
    double beta2 = bseg_cos(angle) * 2.0;
    double beta0 = bseg_cos(angle + two_thirds_pi) * 2.0;
    double beta1 = -(beta0 + beta2);
    double ev0 = q + p * beta0, ev1 = q + p * beta1, ev2 = q + p * beta2;
    if (evals) { evals[0] = ev0 * mc; evals[1] = ev1 * mc; evals[2] = ev2 * mc; }
    double e0[3], e1[3], e2[3];
    if (half_det >= 0.0) {
      bseg_evec0(a00, a01, a02, a11, a12, a22, ev2, e2);
      if (ev2 < ev0 && ev2 < ev1) { out[0] = e2[0]; out[1] = e2[1]; out[2] = e2[2]; return; }
      bseg_evec1(a00, a01, a02, a11, a12, a22, e2, ev1, e1);
      if (ev1 < ev0 && ev1 < ev2) { out[0] = e1[0]; out[1] = e1[1]; out[2] = e1[2]; return; }
      out[0] = e1[1] * e2[2] - e1[2] * e2[1];
      out[1] = e1[2] * e2[0] - e1[0] * e2[2];
      out[2] = e1[0] * e2[1] - e1[1] * e2[0];
    } else {
      bseg_evec0(a00, a01, a02, a11, a12, a22, ev0, e0);
      if (ev0 < ev1 && ev0 < ev2) { out[0] = e0[0]; out[1] = e0[1]; out[2] = e0[2]; return; }
      bseg_evec1(a00, a01, a02, a11, a12, a22, e0, ev1, e1);
      if (ev1 < ev0 && ev1 < ev2) { out[0] = e1[0]; out[1] = e1[1]; out[2] = e1[2]; return; }
      out[0] = e0[1] * e1[2] - e0[2] * e1[1];
      out[1] = e0[2] * e1[0] - e0[0] * e1[2];
      out[2] = e0[0] * e1[1] - e0[1] * e1[0];
    }
  } else {
    /* diagonal matrix: axis of the strictly smallest diagonal entry, else +z */
    if (evals) {
      double d0 = a00 * mc, d1 = a11 * mc, d2 = a22 * mc, t;
      if (d0 > d1) { t = d0; d0 = d1; d1 = t; }
      if (d1 > d2) { t = d1; d1 = d2; d2 = t; }
      if (d0 > d1) { t = d0; d0 = d1; d1 = t; }
      evals[0] = d0; evals[1] = d1; evals[2] = d2;
    }
    if (a00 < a11 && a00 < a22) { out[0] = 1.0; out[1] = 0.0; out[2] = 0.0; }
    else if (a11 < a00 && a11 < a22) { out[0] = 0.0; out[1] = 1.0; out[2] = 0.0; }
    else { out[0] = 0.0; out[1] = 0.0; out[2] = 1.0; }
  }
}

/* Full per-point normal: n_hyb hybrid neighbours (self included) -> oriented unit normal.
 *   n_hyb < 3  -> covariance = identity -> (0,0,1)          (EstimatePerPointCovariances)
 *   zero / NaN-norm result -> (0,0,1)                       (EstimateNormals)
 *   n.z < 0 -> -n                                           (OrientNormalsToAlignWithDirection)
 * curvature (an extension -- the reference computes none): lambda0 / (l0+l1+l2), 0 if undefined.
 */
BSEG_HD void bseg_normal_from_sums(const double sums[9], int n_hyb, double nrm[3], double* curvature)
{
  double cov[6];
  double ev[3];
  if (n_hyb >= 3) {
    bseg_covariance_from_sums(sums, n_hyb, cov);
  } else {
    cov[0] = 1.0; cov[1] = 0.0; cov[2] = 0.0; cov[3] = 1.0; cov[4] = 0.0; cov[5] = 1.0;
  }
  bseg_fast_eigen3x3(cov, nrm, ev);
  double nn = nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2];
  if (nn == 0.0) { /* Eigen's norm()==0 test; NaN falls through unchanged, as upstream */
    nrm[0] = 0.0; nrm[1] = 0.0; nrm[2] = 1.0;
  }
  if (nrm[2] < 0.0) { /* n . (0,0,1) < 0 */
    nrm[0] = nrm[0] * -1.0; nrm[1] = nrm[1] * -1.0; nrm[2] = nrm[2] * -1.0;
  }
  if (curvature) {
    double tr = ev[0] + ev[1] + ev[2];
    *curvature = (tr > 0.0) ? ev[0] / tr : 0.0;
  }
}

#endif /* BSEG_ARITH_H */
