/*
 * bseg_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under buildingsegment_b200/ links, imports or calls it.
 *
 * It restates, in plain C, the reference's per-point segmentation path
 * (citations are into /root/reference/tmc3/):
 *
 *   orc_bbox_shift   buildingSeg::buildingSeg            TMC3.cpp:55-79
 *   orc_knn          KDTreeFlann::SearchKNN / the kNN half of SearchHybrid
 *                                                        my_function.h:63,71-78 (Open3D 0.19.0, un-vendored)
 *   orc_normals      EstimateNormals + OrientNormalsToAlignWithDirection
 *                                                        my_function.h:63-68   (Open3D 0.19.0, un-vendored)
 *   orc_grow         seg_plane::get_planes / Broad       my_function.cpp:180-258
 *   orc_paint        seg_plane::set_plane_color          my_function.cpp:260-275
 *   orc_ground_th    buildingSeg::groundTH               TMC3.cpp:181-198
 *   orc_raster       buildingSeg::compute_gird_picture   TMC3.cpp:127-172
 *   orc_save_image   buildingSeg::save_image (pixels only, no PNG) TMC3.cpp:81-121
 *
 * PARITY PINNING
 *   - orc_grow / orc_paint / orc_ground_th / orc_raster / orc_save_image are pinned: tests compare
 *     them bit-for-bit with the reference's own lines compiled verbatim into oracle/_ref
 *     (oracle/build_ref.sh), doubles included: log() is the platform libm's on both sides (the product
 *     finalises the count channel on the host with the same libm, csrc/raster.cu).
 *   - orc_knn / orc_normals are "parity unpinned": the arithmetic lives in Open3D 0.19.0 /
 *     nanoflann, which are not under /root/reference and not installable here, and the reference
 *     has no tests or golden vectors.  They restate the published algorithms (Open3D
 *     EstimateNormals.cpp, KDTreeFlann.cpp, utility/Eigen.cpp) and are cross-checked against
 *     scipy cKDTree, numpy eigh and an INDEPENDENT numpy transcription of ComputeCovariance /
 *     FastEigen3x3 / ComputeEigenvector0/1 (tests/open3d_restated.py, which does not include the
 *     shared header below) on 10^6 covariances, degenerate families included.  nanoflann's tie order is tree-traversal order
 *     (implementation-defined), so ties are canonicalised to (d^2 ascending, index ascending).
 *
 * Build: gcc -O2 -ffp-contract=off -fwrapv -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../buildingsegment_b200/csrc/bseg_arith.h"

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------------
 * TMC3.cpp:58-72 -- bbox over all points, then subtract the minimum from every point in place.
 * Returns width/height of the raster too (TMC3.cpp:75-76) through wh when non-NULL.
 */
ORC_API int orc_bbox_shift(int32_t* xyz, int64_t n, int32_t mn[3], int32_t mx[3], int bin, int32_t wh[2])
{
  for (int k = 0; k < 3; ++k) {
    mn[k] = INT32_MAX;
    mx[k] = INT32_MIN;
  }
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      int32_t v = xyz[3 * i + k];
      if (v > mx[k]) mx[k] = v;
      if (v < mn[k]) mn[k] = v;
    }
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k)
      xyz[3 * i + k] = (int32_t)((uint32_t)xyz[3 * i + k] - (uint32_t)mn[k]);
  if (wh && n > 0) {
    wh[0] = (mx[0] - mn[0]) / bin + 2;
    wh[1] = (mx[1] - mn[1]) / bin + 2;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Exact k-nearest neighbours, self included, ordered by (d^2, index).  Rows are padded with -1
 * (d2 = -1) when fewer than kq points exist (SearchKNN returns min(k, N) entries).
 * Uniform grid + ring expansion; a ring R guarantees every unseen point has d > R*cell.
 */
typedef struct {
  int64_t d2;
  int32_t idx;
} orc_cand;

static inline int orc_less(int64_t d2a, int32_t ia, int64_t d2b, int32_t ib)
{
  return d2a < d2b || (d2a == d2b && ia < ib);
}

ORC_API int orc_knn(const int32_t* xyz, int64_t n, int kq, int32_t cell, int32_t* idx_out, int64_t* d2_out)
{
  if (kq <= 0 || kq > 256 || cell <= 0)
    return -1;
  if (n == 0)
    return 0;
  int32_t mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      int32_t v = xyz[3 * i + k];
      if (v > mx[k]) mx[k] = v;
      if (v < mn[k]) mn[k] = v;
    }
  int64_t g[3];
  int64_t c = cell;
  for (;;) {
    for (int k = 0; k < 3; ++k)
      g[k] = ((int64_t)mx[k] - mn[k]) / c + 1;
    if (g[0] * g[1] * g[2] <= 8 * n + 4096)
      break;
    c *= 2;
  }
  int64_t ncell = g[0] * g[1] * g[2];
  int64_t* start = (int64_t*)calloc((size_t)ncell + 1, sizeof(int64_t));
  int32_t* order = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  int64_t* cid = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  if (!start || !order || !cid)
    return -2;
  for (int64_t i = 0; i < n; ++i) {
    int64_t cx = ((int64_t)xyz[3 * i] - mn[0]) / c, cy = ((int64_t)xyz[3 * i + 1] - mn[1]) / c,
            cz = ((int64_t)xyz[3 * i + 2] - mn[2]) / c;
    cid[i] = (cz * g[1] + cy) * g[0] + cx;
    start[cid[i] + 1]++;
  }
  for (int64_t k = 0; k < ncell; ++k)
    start[k + 1] += start[k];
  {
    int64_t* fill = (int64_t*)malloc((size_t)ncell * sizeof(int64_t));
    memcpy(fill, start, (size_t)ncell * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i)
      order[fill[cid[i]]++] = (int32_t)i; /* ascending index inside a cell */
    free(fill);
  }
  int64_t maxring = g[0];
  if (g[1] > maxring) maxring = g[1];
  if (g[2] > maxring) maxring = g[2];

#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t s = 0; s < n; ++s) {
    int32_t q = order[s];
    int64_t qx = xyz[3 * q], qy = xyz[3 * q + 1], qz = xyz[3 * q + 2];
    int64_t cx = (qx - mn[0]) / c, cy = (qy - mn[1]) / c, cz = (qz - mn[2]) / c;
    orc_cand best[256];
    int nb = 0;
    for (int64_t R = 0; R <= maxring; ++R) {
      for (int64_t z = cz - R; z <= cz + R; ++z) {
        if (z < 0 || z >= g[2]) continue;
        for (int64_t y = cy - R; y <= cy + R; ++y) {
          if (y < 0 || y >= g[1]) continue;
          int shell_yz = (z == cz - R || z == cz + R || y == cy - R || y == cy + R);
          for (int64_t x = cx - R; x <= cx + R; ++x) {
            if (x < 0 || x >= g[0]) continue;
            if (!shell_yz && x != cx - R && x != cx + R) {
              x = cx + R - 1; /* skip the interior: only the new shell of ring R */
              continue;
            }
            int64_t cc = (z * g[1] + y) * g[0] + x;
            for (int64_t t = start[cc]; t < start[cc + 1]; ++t) {
              int32_t j = order[t];
              int64_t dx = xyz[3 * j] - qx, dy = xyz[3 * j + 1] - qy, dz = xyz[3 * j + 2] - qz;
              int64_t d2 = dx * dx + dy * dy + dz * dz;
              if (nb == kq && !orc_less(d2, j, best[nb - 1].d2, best[nb - 1].idx))
                continue;
              int p = nb < kq ? nb : kq - 1;
              while (p > 0 && orc_less(d2, j, best[p - 1].d2, best[p - 1].idx)) {
                best[p] = best[p - 1];
                --p;
              }
              best[p].d2 = d2;
              best[p].idx = j;
              if (nb < kq) ++nb;
            }
          }
        }
      }
      int64_t reach = R * c;
      if (nb == kq && best[nb - 1].d2 <= reach * reach)
        break;
    }
    for (int k = 0; k < kq; ++k) {
      idx_out[(int64_t)q * kq + k] = k < nb ? best[k].idx : -1;
      if (d2_out)
        d2_out[(int64_t)q * kq + k] = k < nb ? best[k].d2 : -1;
    }
  }
  free(start);
  free(order);
  free(cid);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * my_function.h:63-68.  knn/d2 are rows of the ordered kq-NN (kq >= max_nn).  The hybrid set is
 * the prefix of the first max_nn entries with d^2 < radius^2 (KDTreeFlann::SearchHybrid:
 * knnSearch(max_nn) then lower_bound on radius^2).
 */
ORC_API int orc_normals(const int32_t* xyz, int64_t n, const int32_t* knn, const int64_t* d2, int kq,
                        double radius, int max_nn, double* normals, double* curvature, int32_t* n_hyb_out)
{
  if (kq < max_nn && n > kq)
    return -1;
  double r2 = radius * radius;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double sums[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    int lim = max_nn < kq ? max_nn : kq;
    for (int k = 0; k < lim; ++k) {
      int32_t j = knn[i * kq + k];
      if (j < 0 || !((double)d2[i * kq + k] < r2))
        break;
      double x = (double)xyz[3 * j], y = (double)xyz[3 * j + 1], z = (double)xyz[3 * j + 2];
      sums[0] += x; sums[1] += y; sums[2] += z;
      sums[3] += x * x; sums[4] += x * y; sums[5] += x * z;
      sums[6] += y * y; sums[7] += y * z; sums[8] += z * z;
      ++cnt;
    }
    double cv;
    bseg_normal_from_sums(sums, cnt, &normals[3 * i], &cv);
    if (curvature) curvature[i] = cv;
    if (n_hyb_out) n_hyb_out[i] = cnt;
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * my_function.cpp:180-258 restated: iterative DFS with explicit frames and O(1) running sums.
 * Bit-identical to the reference's recursive, re-summing version (the list only grows and both
 * add left-to-right from zero); pinned against oracle/_ref in tests/test_oracle_ref.py.
 *
 * neigh is N x K (row i = K nearest of i, entry 0 skipped blindly as in :224); -1 entries (only
 * when N < K, where the reference reads out of bounds) are treated as "not accepted".
 *
 * Outputs
 *   plane_idx[N]   the reference's Cloud.planeIdx after get_planes (orphan marks included)
 *   label[N]       0, or the id of the LAST plane whose pointIdx holds the point (what
 *                  set_plane_color paints)
 *   planes         seed / id / normal / centre / CSR offsets + point_idx in pointIdx order,
 *                  duplicates kept.  cap_planes / cap_idx bound the arrays; returns the number
 *                  of planes, or -1 on overflow.
 */
typedef struct {
  int64_t cur, end;
} orc_frame;

ORC_API int64_t orc_grow(const int32_t* xyz, int64_t n, const double* normal, const int32_t* neigh, int K,
                         int th_thickness, int th_count, double th_dot, int32_t* plane_idx, int32_t* label,
                         int32_t* plane_seed, double* plane_normal, int32_t* plane_center,
                         int64_t* plane_off, int32_t* point_idx, int64_t cap_planes, int64_t cap_idx,
                         int64_t* n_steps)
{
  int64_t np = 0, steps = 0;
  int32_t cur_id = 1;
  for (int64_t i = 0; i < n; ++i) {
    plane_idx[i] = -1;
    label[i] = 0;
  }
  int64_t cap_list = 1024, cap_stack = 1024;
  int32_t* list = (int32_t*)malloc((size_t)cap_list * sizeof(int32_t));
  orc_frame* stack = (orc_frame*)malloc((size_t)cap_stack * sizeof(orc_frame));
  plane_off[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (plane_idx[i] != -1)
      continue;
    int64_t len = 0, sp = 0;
    list[len++] = (int32_t)i;
    double mn0 = normal[3 * i], mn1 = normal[3 * i + 1], mn2 = normal[3 * i + 2]; /* cur_normal */
    int32_t mc0 = xyz[3 * i], mc1 = xyz[3 * i + 1], mc2 = xyz[3 * i + 2];          /* cur_center */
    double sn0 = 0.0 + mn0, sn1 = 0.0 + mn1, sn2 = 0.0 + mn2;                      /* running sums */
    uint32_t sc0 = (uint32_t)mc0, sc1 = (uint32_t)mc1, sc2 = (uint32_t)mc2;
    int64_t node = i;
    int depth0 = 1, ok = 1;
    for (;;) {
      /* ---- one Broad(node) call: my_function.cpp:221-250 ---- */
      ++steps;
      int64_t s0 = len;
      for (int j = 1; j < K; ++j) {
        int32_t id = neigh[node * K + j];
        if (id < 0)
          continue;
        if (plane_idx[id] <= 0) {
          int32_t p0 = (int32_t)((uint32_t)xyz[3 * id] - (uint32_t)mc0);
          int32_t p1 = (int32_t)((uint32_t)xyz[3 * id + 1] - (uint32_t)mc1);
          int32_t p2 = (int32_t)((uint32_t)xyz[3 * id + 2] - (uint32_t)mc2);
          double dist = fabs(p0 * mn0 + p1 * mn1 + p2 * mn2);
          if (dist <= (double)th_thickness &&
              mn0 * normal[3 * id] + mn1 * normal[3 * id + 1] + mn2 * normal[3 * id + 2] >= th_dot) {
            if (len == cap_list) {
              cap_list *= 2;
              list = (int32_t*)realloc(list, (size_t)cap_list * sizeof(int32_t));
            }
            list[len++] = id;
            plane_idx[id] = cur_id;
            sn0 += normal[3 * id]; sn1 += normal[3 * id + 1]; sn2 += normal[3 * id + 2];
            sc0 += (uint32_t)xyz[3 * id]; sc1 += (uint32_t)xyz[3 * id + 1]; sc2 += (uint32_t)xyz[3 * id + 2];
          }
        }
      }
      if (depth0 && (len - s0) < K - 1) {
        ok = 0; /* :238-239 -- marks stay (orphans) */
        break;
      }
      depth0 = 0;
      {
        double nn = bseg_sqrt((sn0 * sn0) + (sn1 * sn1) + (sn2 * sn2)); /* :249, PCCMath.h:95-98 */
        mn0 = sn0 / nn; mn1 = sn1 / nn; mn2 = sn2 / nn;
        uint64_t dv = (uint64_t)len; /* :250 -- int /= size_t goes through uint64 */
        mc0 = (int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)sc0) / dv);
        mc1 = (int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)sc1) / dv);
        mc2 = (int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)sc2) / dv);
      }
      if (len > s0) {
        if (sp == cap_stack) {
          cap_stack *= 2;
          stack = (orc_frame*)realloc(stack, (size_t)cap_stack * sizeof(orc_frame));
        }
        stack[sp].cur = s0;
        stack[sp].end = len;
        ++sp;
      }
      /* ---- next call in DFS pre-order: :252-255 ---- */
      while (sp > 0 && stack[sp - 1].cur == stack[sp - 1].end)
        --sp;
      if (sp == 0)
        break;
      node = list[stack[sp - 1].cur++];
    }
    if (!ok)
      continue;
    if (len > th_count) { /* :199 (size_t > int) */
      if (np >= cap_planes || plane_off[np] + len > cap_idx) {
        free(list);
        free(stack);
        return -1;
      }
      plane_seed[np] = (int32_t)i;
      plane_normal[3 * np] = mn0; plane_normal[3 * np + 1] = mn1; plane_normal[3 * np + 2] = mn2;
      plane_center[3 * np] = mc0; plane_center[3 * np + 1] = mc1; plane_center[3 * np + 2] = mc2;
      memcpy(point_idx + plane_off[np], list, (size_t)len * sizeof(int32_t));
      plane_off[np + 1] = plane_off[np] + len;
      for (int64_t t = 0; t < len; ++t)
        label[list[t]] = cur_id;
      ++np;
      ++cur_id;
    } else {
      for (int64_t t = 0; t < len; ++t)
        plane_idx[list[t]] = -1; /* :203-209 */
    }
  }
  free(list);
  free(stack);
  if (n_steps) *n_steps = steps;
  return np;
}

/* my_function.cpp:260-275 -- colours: zero everything, then plane p paints its pointIdx with
 * plane_rgb[p] (the host draws 55+rand()%200 three times per plane, left to right). */
ORC_API int orc_paint(int64_t n, int64_t n_planes, const int64_t* plane_off, const int32_t* point_idx,
                      const uint16_t* plane_rgb, uint16_t* colors)
{
  memset(colors, 0, (size_t)n * 3 * sizeof(uint16_t));
  for (int64_t p = 0; p < n_planes; ++p)
    for (int64_t t = plane_off[p]; t < plane_off[p + 1]; ++t) {
      int32_t id = point_idx[t];
      colors[3 * id] = plane_rgb[3 * p];
      colors[3 * id + 1] = plane_rgb[3 * p + 1];
      colors[3 * id + 2] = plane_rgb[3 * p + 2];
    }
  return 0;
}

/* TMC3.cpp:181-198 on the already shifted cloud (zmin = 0, zmax = mx-mn). */
ORC_API double orc_ground_th(const int32_t* xyz, int64_t n, int32_t zext, int bin_height)
{
  int64_t nb = zext / bin_height + 1;
  int32_t* h = (int32_t*)calloc((size_t)nb, sizeof(int32_t));
  int TH = (int)(n / 2);
  for (int64_t i = 0; i < n; ++i)
    h[xyz[3 * i + 2] / bin_height]++;
  int total = 0;
  int64_t i;
  for (i = 0; i < nb; ++i) {
    total += h[i];
    if (total > TH)
      break;
  }
  free(h);
  return (double)((int)i * bin_height);
}

/* TMC3.cpp:127-164.  image is W*H*3 doubles, zero-initialised here. */
static int orc_raster_impl(const int32_t* xyz, int64_t n, double th, int bin, double bias, int W, int H, double* image);

ORC_API int orc_raster(const int32_t* xyz, int64_t n, int32_t zext, int bin, int bin_height, double bias,
                       int W, int H, double* image)
{
  return orc_raster_impl(xyz, n, orc_ground_th(xyz, n, zext, bin_height), bin, bias, W, H, image);
}

/* the same with the ground threshold given (a slab of a tile uses the tile's: include/bseg.h bseg_raster_device) */
ORC_API int orc_raster_th(const int32_t* xyz, int64_t n, double th, int bin, double bias, int W, int H, double* image)
{
  return orc_raster_impl(xyz, n, th, bin, bias, W, H, image);
}

static int orc_raster_impl2(const int32_t* xyz, int64_t n, double th, int bin, double bias, int W, int H, double* image,
                            int sums_only);
static int orc_raster_impl(const int32_t* xyz, int64_t n, double th, int bin, double bias, int W, int H, double* image)
{
  return orc_raster_impl2(xyz, n, th, bin, bias, W, H, image, 0);
}

/* channel 1 left as the weight sums (before TMC3.cpp:159-164): what a slab hands to the host half of the product */
ORC_API int orc_raster_th_sums(const int32_t* xyz, int64_t n, double th, int bin, int W, int H, double* image)
{
  return orc_raster_impl2(xyz, n, th, bin, 0.0, W, H, image, 1);
}

static int orc_raster_impl2(const int32_t* xyz, int64_t n, double th, int bin, double bias, int W, int H, double* image,
                            int sums_only)
{
  memset(image, 0, (size_t)W * H * 3 * sizeof(double));
  for (int64_t i = 0; i < n; ++i) {
    int32_t p0 = xyz[3 * i], p1 = xyz[3 * i + 1], p2 = xyz[3 * i + 2];
    int x = p0 / bin, y = p1 / bin;
    if (p2 < th)
      continue;
    for (int xi = 0; xi < 2; ++xi)
      for (int yi = 0; yi < 2; ++yi) {
        double w = 1.0 * p0 / bin - x;
        double h = 1.0 * p1 / bin - y;
        double s = ((xi == 1) ? w : (1 - w)) * ((yi == 1) ? h : (1 - h));
        size_t px = ((size_t)(y + yi) * W + (x + xi)) * 3;
        image[px + 1] += s;
        image[px + 0] += s * p2;
      }
  }
  for (size_t px = 0; px < (size_t)W * H; ++px)
    if (image[3 * px + 1] != 0)
      image[3 * px] = image[3 * px] / image[3 * px + 1];
  if (sums_only)
    return 0;
  for (size_t px = 0; px < (size_t)W * H; ++px) {
    image[3 * px + 1] = log(image[3 * px + 1] + 1); /* the platform's libm, as the reference's std::log (TMC3.cpp:161) */
    if (image[3 * px + 1] != 0)
      image[3 * px + 1] += bias;
  }
  return 0;
}

/* TMC3.cpp:81-119 minus the PNG encode: three W*H*3 uint8 images (A: ch0->byte0,
 * B: ch1->byte1, C: ch2->byte1). */
ORC_API int orc_save_image(const double* image, int W, int H, uint8_t* imgA, uint8_t* imgB, uint8_t* imgC,
                           double max_out[3])
{
  double mx[3] = {0, 0, 0};
  size_t np = (size_t)W * H;
  for (size_t px = 0; px < np; ++px)
    for (int ch = 0; ch < 3; ++ch)
      if (mx[ch] < image[3 * px + ch])
        mx[ch] = image[3 * px + ch];
  memset(imgA, 0, np * 3);
  memset(imgB, 0, np * 3);
  memset(imgC, 0, np * 3);
  for (size_t px = 0; px < np; ++px) {
    if (mx[0] != 0) imgA[3 * px + 0] = (uint8_t)(255.0 * (1.0 * image[3 * px + 0] / mx[0]));
    if (mx[1] != 0) imgB[3 * px + 1] = (uint8_t)(255.0 * (1.0 * image[3 * px + 1] / mx[1]));
    if (mx[2] != 0) imgC[3 * px + 1] = (uint8_t)(255.0 * (1.0 * image[3 * px + 2] / mx[2]));
  }
  if (max_out) {
    max_out[0] = mx[0]; max_out[1] = mx[1]; max_out[2] = mx[2];
  }
  return 0;
}

/* helpers exported so tests can probe the shared arithmetic from the host side */
ORC_API double orc_acos(double x) { return bseg_acos(x); }
ORC_API double orc_cos(double x) { return bseg_cos(x); }
ORC_API double orc_log(double x) { return bseg_log(x); }
ORC_API void orc_eigen(const double cov[6], double out[3], double ev[3]) { bseg_fast_eigen3x3(cov, out, ev); }
/* vector forms for the independent-restatement tests: which = 0 acos, 1 cos (the shared polynomial kernels) */
ORC_API void orc_trig_vec(const double* x, int64_t n, int which, double* out)
{
  for (int64_t i = 0; i < n; ++i) out[i] = which == 0 ? bseg_acos(x[i]) : bseg_cos(x[i]);
}
ORC_API void orc_eigen_vec(const double* cov6, int64_t n, double* out3)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) bseg_fast_eigen3x3(cov6 + 6 * i, out3 + 3 * i, NULL);
}
/* covariance -> oriented normal exactly as orc_normals finishes a point (zero rule, +z orientation) */
ORC_API void orc_normal_from_sums_vec(const double* sums9, const int32_t* cnt, int64_t n, double* out3)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) bseg_normal_from_sums(sums9 + 9 * i, cnt[i], out3 + 3 * i, NULL);
}
ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
