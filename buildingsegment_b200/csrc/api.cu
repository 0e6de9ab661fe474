// api.cu -- the extern "C" boundary of libbseg.so (include/bseg.h): context lifetime, device
// memory, host<->device marshalling and per-stage timing.  Kernels live in bin/knn/grow/raster.cu.
//
// There is no CPU path: bseg_create fails with BSEG_E_NODEVICE when no CUDA device is usable.
#include <stdarg.h>

#include <new>

#include "common.cuh"

static thread_local char g_err[512] = "";

int bseg_fail(bseg_ctx* c, int code, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (c)
    memcpy(c->err, g_err, sizeof(c->err));
  return code;
}

int dev_ensure(bseg_ctx* c, DevBuf& b, size_t bytes)
{
  if (bytes == 0)
    bytes = 16;
  if (b.cap >= bytes)
    return 0;
  if (b.p) {
    cudaStreamSynchronize(c->stream);
    cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 16 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    cudaGetLastError();
    return bseg_fail(c, BSEG_E_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  b.cap = want;
  return 0;
}

void dev_free(DevBuf& b)
{
  if (b.p)
    cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

extern "C" {

BSEG_API const char* bseg_version(void) { return "bseg-b200 0.1 (sm_100a)"; }

BSEG_API const char* bseg_last_error(const bseg_ctx* ctx) { return ctx ? ctx->err : g_err; }

BSEG_API int bseg_default_params(bseg_params* p)
{
  if (!p)
    return BSEG_E_ARG;
  memset(p, 0, sizeof(*p));
  p->K = 15;
  p->max_nn = 50;
  p->radius = 100.0;
  p->th_thickness = 300;
  p->th_point_count = 400;
  p->th_dot = 0.88;
  p->bin = 100;
  p->bin_height = 1000;
  p->count_bias = 20.0;
  p->cell = 0;
  p->grow_mode = 0;
  p->grow_radius = 0.0;
  return 0;
}

BSEG_API int bseg_create(bseg_ctx** out, int device)
{
  if (!out)
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_create: out == NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return bseg_fail(nullptr, BSEG_E_NODEVICE, "no CUDA device (%s); libbseg has no CPU path",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= ndev)
    return bseg_fail(nullptr, BSEG_E_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  bseg_ctx* c = new (std::nothrow) bseg_ctx();
  if (!c)
    return bseg_fail(nullptr, BSEG_E_NOMEM, "out of host memory");
  c->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    int rc = bseg_fail(nullptr, BSEG_E_CUDA, "device init failed: %s", cudaGetErrorString(e));
    delete c;
    return rc;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
    c->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
      int rc = bseg_fail(nullptr, BSEG_E_NODEVICE, "device %d is sm_%d%d; libbseg is built for sm_100a only", device,
                         prop.major, prop.minor);
      cudaStreamDestroy(c->stream);
      delete c;
      return rc;
    }
  }
  for (int i = 0; i < EV_COUNT; ++i) {
    cudaEventCreate(&c->ev[i]);
    cudaEventCreate(&c->ev_end[i]);
  }
  memset(&c->tm, 0, sizeof(c->tm));
  *out = c;
  return 0;
}

BSEG_API void bseg_destroy(bseg_ctx* c)
{
  if (!c)
    return;
  cudaSetDevice(c->device);
  raster_host_join(c);
  cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  DevBuf* all[] = {&c->xyz_raw, &c->minmax, &c->keys[0], &c->keys[1], &c->vals[0], &c->vals[1], &c->sort_cnt,
                   &c->scan_tmp, &c->pts, &c->inv, &c->flags, &c->cell_key, &c->cell_start, &c->hash_keys,
                   &c->hash_vals, &c->cell_key2, &c->cell_start2, &c->hash_keys2, &c->hash_vals2, &c->nbr, &c->nrm, &c->curv, &c->worklist, &c->counters, &c->out_tmp, &c->x_neigh, &c->x_normals,
                   &c->g_state, &c->g_res, &c->g_spec, &c->g_pool, &c->g_planes, &c->g_tx, &c->g_queue, &c->g_rowdup, &c->g_marklog,
                   &c->g_stack, &c->g_label, &c->g_pidx, &c->g_pts_raw, &c->g_nbr_masked, &c->r_hist, &c->r_image, &c->r_png, &c->r_pix, &c->r_cnt};
  for (DevBuf* b : all)
    dev_free(*b);
  for (int i = 0; i < EV_COUNT; ++i) {
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->ev_end[i]) cudaEventDestroy(c->ev_end[i]);
  }
  for (auto& e : c->grow_ev)
    if (e) cudaEventDestroy(e);
  if (c->grow_pinned) cudaFreeHost(c->grow_pinned);
  if (c->grow_fork) cudaEventDestroy(c->grow_fork);
  for (auto& e : c->grow_join)
    if (e) cudaEventDestroy(e);
  if (c->grow_hi) cudaStreamDestroy(c->grow_hi);
  if (c->grow_lo) cudaStreamDestroy(c->grow_lo);
  if (c->grow_pf) cudaStreamDestroy(c->grow_pf);
  if (c->out_ready) cudaEventDestroy(c->out_ready);
  if (c->out_stream) cudaStreamDestroy(c->out_stream);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->h_cnt) cudaFreeHost(c->h_cnt);
  if (c->raster_done) cudaEventDestroy(c->raster_done);
  if (c->raster_copied) cudaEventDestroy(c->raster_copied);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  grow_host_free(c);
  cudaStreamDestroy(c->stream);
  delete c;
}

static int check_ctx(bseg_ctx* c)
{
  if (!c)
    return bseg_fail(nullptr, BSEG_E_ARG, "ctx == NULL");
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess)
    return bseg_fail(c, BSEG_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return 0;
}

static int set_points_common(bseg_ctx* c, int64_t n, int32_t out_min[3], int32_t out_max[3])
{
  c->n = n;
  c->n_owned = n;
  c->have_points = true;
  raster_host_join(c);
  c->have_knn = c->have_grow = c->have_bin = false;
  RC_CHECK(stage_bbox_shift(c));
  for (int k = 0; k < 3; ++k) {
    if (out_min) out_min[k] = c->mn[k];
    if (out_max) out_max[k] = c->mx[k];
    if (n > 0 && (int64_t)c->mx[k] - (int64_t)c->mn[k] >= BSEG_MAX_COORD) {
      c->have_points = false;
      return bseg_fail(c, BSEG_E_ARG, "extent of axis %d is %lld units; the supported maximum is %d", k,
                       (long long)c->mx[k] - c->mn[k], BSEG_MAX_COORD - 1);
    }
  }
  return 0;
}

BSEG_API int bseg_set_points(bseg_ctx* c, const int32_t* xyz_aos, int64_t n, int32_t out_min[3], int32_t out_max[3],
                             int32_t* xyz_shifted_out)
{
  RC_CHECK(check_ctx(c));
  if (n < 0 || n >= ((int64_t)1 << 31) - 64 || (n > 0 && !xyz_aos))
    return bseg_fail(c, BSEG_E_ARG, "bseg_set_points: bad n (%lld) or NULL cloud", (long long)n);
  RC_CHECK(dev_ensure(c, c->xyz_raw, (size_t)n * 12));
  STAGE_BEGIN(c, EV_H2D);
  if (n > 0)
    CU_CHECK(c, cudaMemcpyAsync(c->xyz_raw.p, xyz_aos, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
  STAGE_END(c, EV_H2D);
  RC_CHECK(set_points_common(c, n, out_min, out_max));
  if (xyz_shifted_out && n > 0) {
    CU_CHECK(c, cudaMemcpyAsync(xyz_shifted_out, c->xyz_raw.p, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  return 0;
}

BSEG_API int bseg_set_points_device(bseg_ctx* c, const int32_t* d_xyz_aos, int64_t n, int32_t out_min[3],
                                    int32_t out_max[3])
{
  RC_CHECK(check_ctx(c));
  if (n < 0 || n >= ((int64_t)1 << 31) - 64 || (n > 0 && !d_xyz_aos))
    return bseg_fail(c, BSEG_E_ARG, "bseg_set_points_device: bad n (%lld) or NULL cloud", (long long)n);
  RC_CHECK(dev_ensure(c, c->xyz_raw, (size_t)n * 12));
  if (n > 0)
    CU_CHECK(c, cudaMemcpyAsync(c->xyz_raw.p, d_xyz_aos, (size_t)n * 12, cudaMemcpyDeviceToDevice, c->stream));
  return set_points_common(c, n, out_min, out_max);
}

BSEG_API int bseg_set_owned(bseg_ctx* c, int64_t n_owned)
{
  RC_CHECK(check_ctx(c));
  if (!c->have_points || n_owned < 0 || n_owned > c->n)
    return bseg_fail(c, BSEG_E_ARG, "bseg_set_owned: n_owned %lld outside 0..%lld", (long long)n_owned, (long long)c->n);
  c->n_owned = n_owned;
  return 0;
}

BSEG_API int bseg_set_origin(bseg_ctx* c, const int32_t* origin)
{
  RC_CHECK(check_ctx(c));
  c->have_origin = origin != nullptr;
  for (int k = 0; k < 3; ++k) c->origin[k] = origin ? origin[k] : 0;
  return 0;
}

BSEG_API int bseg_set_grow_offset(bseg_ctx* c, const int32_t* offset)
{
  RC_CHECK(check_ctx(c));
  for (int k = 0; k < 3; ++k) c->grow_off[k] = offset ? offset[k] : 0;
  c->have_grow = false;
  return 0;
}

BSEG_API int bseg_device_results(bseg_ctx* c, const int32_t** d_label, const int32_t** d_plane_idx,
                                 const int32_t** d_xyz_shifted)
{
  RC_CHECK(check_ctx(c));
  if (!c->have_points)
    return bseg_fail(c, BSEG_E_STATE, "bseg_device_results: no cloud");
  if ((d_label || d_plane_idx) && !c->have_grow)
    return bseg_fail(c, BSEG_E_STATE, "bseg_device_results: the grower has not run");
  if (d_label) *d_label = dptr<int32_t>(c->g_label);
  if (d_plane_idx) *d_plane_idx = dptr<int32_t>(c->g_pidx);
  if (d_xyz_shifted) *d_xyz_shifted = dptr<int32_t>(c->xyz_raw);
  return 0;
}

BSEG_API int bseg_halo_check(bseg_ctx* c, int32_t x_lo, int32_t x_hi, int32_t halo, int64_t* n_unresolved)
{
  RC_CHECK(check_ctx(c));
  if (!c->have_knn || !n_unresolved)
    return bseg_fail(c, BSEG_E_STATE, "bseg_halo_check: run bseg_knn_normals first");
  return stage_halo_check(c, x_lo, x_hi, halo, n_unresolved);
}

static int check_params(bseg_ctx* c, const bseg_params* p)
{
  if (!p)
    return bseg_fail(c, BSEG_E_ARG, "params == NULL");
  if (p->K < 2 || p->K > BSEG_MAX_K)
    return bseg_fail(c, BSEG_E_ARG, "K = %d outside 2..%d", p->K, BSEG_MAX_K);
  if (p->max_nn < 1 || p->max_nn > 4096)
    return bseg_fail(c, BSEG_E_ARG, "max_nn = %d outside 1..4096", p->max_nn);
  if (!(p->radius > 0.0) || p->radius > (double)BSEG_MAX_CELL)
    return bseg_fail(c, BSEG_E_ARG, "radius = %g outside (0, %d]", p->radius, BSEG_MAX_CELL);
  if (p->cell < 0 || p->cell > BSEG_MAX_CELL || (p->cell > 0 && (double)p->cell < p->radius))
    return bseg_fail(c, BSEG_E_ARG, "cell = %d must be 0 (auto) or in [radius, %d]", p->cell, BSEG_MAX_CELL);
  if (p->bin < 1 || p->bin_height < 1)
    return bseg_fail(c, BSEG_E_ARG, "bin / bin_height must be positive");
  if (!(p->grow_radius >= 0.0) || p->grow_radius > 3.0e4)
    return bseg_fail(c, BSEG_E_ARG, "grow_radius = %g outside [0, 30000]", p->grow_radius);
  if (p->th_thickness < 0 || p->th_point_count < 0)
    return bseg_fail(c, BSEG_E_ARG, "negative threshold");
  return 0;
}

BSEG_API int bseg_knn_normals(bseg_ctx* c, const bseg_params* p, int32_t* neigh_NxK, double* normals_Nx3,
                              double* curvature_N)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_points)
    return bseg_fail(c, BSEG_E_STATE, "bseg_knn_normals before bseg_set_points");
  RC_CHECK(stage_bin(c, p));
  c->have_bin = true;
  RC_CHECK(stage_knn(c, p));
  c->have_knn = true;
  c->have_grow = false;
  if (neigh_NxK || normals_Nx3 || curvature_N)
    RC_CHECK(stage_export_knn(c, p, neigh_NxK, normals_Nx3, curvature_N));
  return 0;
}

BSEG_API int bseg_override_neigh_normals(bseg_ctx* c, const bseg_params* p, const int32_t* neigh_NxK,
                                         const double* normals_Nx3)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_points)
    return bseg_fail(c, BSEG_E_STATE, "bseg_override_neigh_normals before bseg_set_points");
  if (!c->have_knn || c->K != p->K) {
    RC_CHECK(stage_bin(c, p));
    c->have_bin = true;
    RC_CHECK(stage_knn(c, p));
    c->have_knn = true;
  }
  c->have_grow = false;
  return stage_override(c, p, neigh_NxK, normals_Nx3, false);
}

BSEG_API int bseg_knn_device_results(bseg_ctx* c, const bseg_params* p, const int32_t** d_neigh_NxK,
                                     const double** d_normals_Nx3)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_knn || c->K != p->K)
    return bseg_fail(c, BSEG_E_STATE, "bseg_knn_device_results: run bseg_knn_normals / bseg_run_device(KNN) with this K first");
  return stage_export_knn_device(c, p, d_neigh_NxK, d_normals_Nx3);
}

BSEG_API int bseg_import_neigh_normals_device(bseg_ctx* c, const bseg_params* p, const int32_t* d_neigh_NxK,
                                              const double* d_normals_Nx3)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_points)
    return bseg_fail(c, BSEG_E_STATE, "bseg_import_neigh_normals_device before bseg_set_points");
  if (!d_neigh_NxK || !d_normals_Nx3)
    return bseg_fail(c, BSEG_E_ARG, "bseg_import_neigh_normals_device: both arrays are required");
  if (!c->have_bin || c->K != p->K) {
    RC_CHECK(stage_bin(c, p));
    c->have_bin = true;
  }
  RC_CHECK(stage_alloc_knn_outputs(c, p));
  c->have_knn = true;
  c->have_grow = false;
  return stage_override(c, p, d_neigh_NxK, d_normals_Nx3, true);
}

BSEG_API int bseg_grow_planes(bseg_ctx* c, const bseg_params* p, int32_t* plane_idx_N, int32_t* label_N,
                              int32_t* n_planes)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_knn)
    return bseg_fail(c, BSEG_E_STATE, "bseg_grow_planes before bseg_knn_normals");
  if (c->K != p->K)
    return bseg_fail(c, BSEG_E_STATE, "K changed between bseg_knn_normals (%d) and bseg_grow_planes (%d)", c->K, p->K);
  RC_CHECK(stage_grow(c, p));
  c->have_grow = true;
  if (n_planes)
    *n_planes = c->n_planes;
  if (plane_idx_N || label_N)
    RC_CHECK(stage_export_grow(c, plane_idx_N, label_N));
  return 0;
}

BSEG_API int bseg_get_planes(bseg_ctx* c, int32_t* seeds_P, double* normals_Px3, int32_t* centers_Px3,
                             int64_t* offsets_Pp1, int32_t* point_idx)
{
  RC_CHECK(check_ctx(c));
  if (!c->have_grow)
    return bseg_fail(c, BSEG_E_STATE, "bseg_get_planes before bseg_grow_planes");
  return stage_get_planes(c, seeds_P, normals_Px3, centers_Px3, offsets_Pp1, point_idx);
}

BSEG_API int bseg_paint(bseg_ctx* c, const int32_t* plane_ids_Q, int32_t n_listed, const uint16_t* plane_rgb_Qx3,
                        uint16_t* colors_Nx3)
{
  RC_CHECK(check_ctx(c));
  if (!c->have_grow)
    return bseg_fail(c, BSEG_E_STATE, "bseg_paint before bseg_grow_planes");
  if (!colors_Nx3 || n_listed < 0 || (n_listed > 0 && !plane_rgb_Qx3))
    return bseg_fail(c, BSEG_E_ARG, "bseg_paint: NULL buffer or negative count");
  if (!plane_ids_Q && n_listed != c->n_planes)
    return bseg_fail(c, BSEG_E_ARG, "bseg_paint: %d colours for %d planes (pass plane ids to paint a subset)", n_listed,
                     c->n_planes);
  if (plane_ids_Q)
    for (int32_t q = 0; q < n_listed; ++q)
      if (plane_ids_Q[q] < 1 || plane_ids_Q[q] > c->n_planes)
        return bseg_fail(c, BSEG_E_ARG, "bseg_paint: plane id %d outside 1..%d", plane_ids_Q[q], c->n_planes);
  return stage_paint(c, plane_ids_Q, n_listed, plane_rgb_Qx3, colors_Nx3);
}

BSEG_API int bseg_plane_classes(bseg_ctx* c, double facade_max_nz, double roof_min_nz, double ground_z,
                                double* equations_Px4, uint8_t* plane_class_P, uint8_t* point_class_N)
{
  RC_CHECK(check_ctx(c));
  if (!c->have_grow)
    return bseg_fail(c, BSEG_E_STATE, "bseg_plane_classes before bseg_grow_planes");
  if (!(facade_max_nz >= 0.0) || !(roof_min_nz >= facade_max_nz))
    return bseg_fail(c, BSEG_E_ARG, "bseg_plane_classes: need 0 <= facade_max_nz <= roof_min_nz");
  return stage_plane_classes(c, facade_max_nz, roof_min_nz, ground_z, equations_Px4, plane_class_P, point_class_N);
}

BSEG_API int bseg_contour_mask(bseg_ctx* c, const uint8_t* pixels, int32_t w, int32_t h, int32_t comp, int32_t channel,
                               int32_t thresh, int32_t iterations, uint8_t* mask_out)
{
  RC_CHECK(check_ctx(c));
  if (!pixels || !mask_out || w < 0 || h < 0 || comp < 1 || comp > 4 || channel < 0 || channel >= comp || iterations < 0 ||
      iterations > 64)
    return bseg_fail(c, BSEG_E_ARG, "bseg_contour_mask: bad arguments");
  return stage_contour_mask(c, pixels, w, h, comp, channel, thresh, iterations, mask_out);
}

BSEG_API int bseg_find_contours(const uint8_t* mask, int32_t w, int32_t h, int32_t simple, int32_t* points_xy, int64_t cap_points,
                                int64_t* offsets, int64_t cap_contours, int64_t* n_contours, int64_t* n_points)
{
  if (!mask || w < 0 || h < 0 || !n_contours || !n_points)
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_find_contours: bad arguments");
  std::vector<int32_t> pts;
  std::vector<int64_t> off;
  contour_find_host(mask, w, h, simple != 0, pts, off);
  *n_contours = (int64_t)off.size() - 1;
  *n_points = (int64_t)pts.size() / 2;
  if (points_xy || offsets) {
    if (!points_xy || !offsets || cap_points < *n_points || cap_contours < *n_contours)
      return bseg_fail(nullptr, BSEG_E_CAPACITY, "bseg_find_contours: %lld contours / %lld points do not fit the buffers",
                       (long long)*n_contours, (long long)*n_points);
    if (!pts.empty()) memcpy(points_xy, pts.data(), pts.size() * 4);
    memcpy(offsets, off.data(), off.size() * 8);
  }
  return 0;
}

BSEG_API int bseg_contour_measure(const int32_t* points_xy, int64_t n, double* area, double* perimeter)
{
  if ((n > 0 && !points_xy) || n < 0)
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_contour_measure: bad arguments");
  if (area) *area = contour_area(points_xy, n);
  if (perimeter) *perimeter = contour_perimeter(points_xy, n);
  return 0;
}

BSEG_API int bseg_draw_contour(uint8_t* image, int32_t w, int32_t h, int32_t comp, const int32_t* points_xy, int64_t n,
                               const uint8_t* color3)
{
  if (!image || w < 0 || h < 0 || comp < 1 || comp > 4 || (n > 0 && !points_xy) || !color3)
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_draw_contour: bad arguments");
  contour_draw(image, w, h, comp, points_xy, n, color3);
  return 0;
}

BSEG_API int bseg_raster_size(bseg_ctx* c, const bseg_params* p, int32_t* W, int32_t* H)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_points)
    return bseg_fail(c, BSEG_E_STATE, "bseg_raster_size before bseg_set_points");
  return stage_raster_size(c, p, W, H);
}

BSEG_API int bseg_raster_device(bseg_ctx* c, const bseg_params* p, const double* ground_th, const double** d_image,
                                int32_t* W, int32_t* H)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_points)
    return bseg_fail(c, BSEG_E_STATE, "bseg_raster_device before bseg_set_points");
  RC_CHECK(stage_raster(c, p, nullptr, nullptr, nullptr, nullptr, nullptr, RASTER_DEVICE_ONLY, ground_th));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  if (d_image) *d_image = c->n > 0 ? dptr<double>(c->r_image) : nullptr;
  if (W) *W = c->rW;
  if (H) *H = c->rH;
  return 0;
}

BSEG_API int bseg_label_raster(bseg_ctx* c, const bseg_params* p, const uint16_t* plane_rgb_Px3, int32_t* label_WxH,
                               uint8_t* rgb_WxHx3)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_grow)
    return bseg_fail(c, BSEG_E_STATE, "bseg_label_raster before bseg_grow_planes");
  return stage_label_raster(c, p, plane_rgb_Px3, label_WxH, rgb_WxHx3);
}

BSEG_API int bseg_raster(bseg_ctx* c, const bseg_params* p, double* image_WxHx3, uint8_t* png_a, uint8_t* png_b,
                         uint8_t* png_c, double* ground_th)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_points)
    return bseg_fail(c, BSEG_E_STATE, "bseg_raster before bseg_set_points");
  return stage_raster(c, p, image_WxHx3, png_a, png_b, png_c, ground_th, RASTER_SYNC);
}

BSEG_API int bseg_run_device(bseg_ctx* c, const bseg_params* p, int stages)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  if (!c->have_points)
    return bseg_fail(c, BSEG_E_STATE, "bseg_run_device before bseg_set_points");
  if (stages & BSEG_RUN_KNN) {
    RC_CHECK(stage_bin(c, p));
    c->have_bin = true;
    RC_CHECK(stage_knn(c, p));
    c->have_knn = true;
    c->have_grow = false;
  }
  // the raster does not depend on the planes: its device pass goes first and the host half of its count channel
  // (libm log, raster.cu) runs on a worker thread while the grower occupies the GPU
  if (stages & BSEG_RUN_RASTER)
    RC_CHECK(stage_raster(c, p, nullptr, nullptr, nullptr, nullptr, nullptr, RASTER_ASYNC));
  if (stages & BSEG_RUN_GROW) {
    if (!c->have_knn) {
      raster_host_join(c);
      return bseg_fail(c, BSEG_E_STATE, "grow stage requested before knn");
    }
    const int rc = stage_grow(c, p);
    if (rc != 0) {
      raster_host_join(c);
      return rc;
    }
    c->have_grow = true;
  }
  RC_CHECK(raster_host_join(c));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

BSEG_API int bseg_segment_host(bseg_ctx* c, const bseg_params* p, const int32_t* xyz_aos, int64_t n,
                               int32_t* xyz_shifted_out, int32_t* label_N, int32_t* n_planes, uint8_t* png_a,
                               uint8_t* png_b, int32_t* W, int32_t* H)
{
  RC_CHECK(check_ctx(c));
  RC_CHECK(check_params(c, p));
  int32_t mn[3], mx[3];
  RC_CHECK(bseg_set_points(c, xyz_aos, n, mn, mx, nullptr));
  // the shifted cloud (TMC3.cpp:71 moves the caller's copy too) goes back on its own stream, beside the kNN: nothing
  // writes xyz_raw after the shift
  bool out_pending = false;
  if (xyz_shifted_out && n > 0) {
    if (!c->out_stream) CU_CHECK(c, cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
    if (!c->out_ready) CU_CHECK(c, cudaEventCreateWithFlags(&c->out_ready, cudaEventDisableTiming));
    CU_CHECK(c, cudaEventRecord(c->out_ready, c->stream));
    CU_CHECK(c, cudaStreamWaitEvent(c->out_stream, c->out_ready, 0));
    CU_CHECK(c, cudaMemcpyAsync(xyz_shifted_out, c->xyz_raw.p, (size_t)n * 12, cudaMemcpyDeviceToHost, c->out_stream));
    out_pending = true;
  }
  int rc = bseg_knn_normals(c, p, nullptr, nullptr, nullptr);
  if (rc == 0 && (png_a || png_b || W || H)) {
    rc = stage_raster_size(c, p, W, H);
    if (rc == 0 && (png_a || png_b))  // device pass now, host half (count channel, image B) overlapped with the grower
      rc = stage_raster(c, p, nullptr, png_a, png_b, nullptr, nullptr, RASTER_ASYNC);
  }
  if (rc == 0) rc = bseg_grow_planes(c, p, nullptr, label_N, n_planes);
  const int rc_join = raster_host_join(c);
  if (out_pending) cudaStreamSynchronize(c->out_stream);  // (also on an error path: the caller's buffer must be quiet)
  RC_CHECK(rc);
  RC_CHECK(rc_join);
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

BSEG_API int bseg_count_channel(double* values, int64_t n, double bias, double* max_out)
{
  if (n < 0 || (n > 0 && !values))
    return bseg_fail(nullptr, BSEG_E_ARG, "bseg_count_channel: bad arguments");
  const double m = bseg_count_channel_host(values, n, bias);
  if (max_out) *max_out = m;
  return 0;
}

BSEG_API int bseg_get_timings(bseg_ctx* c, bseg_timings* out)
{
  RC_CHECK(check_ctx(c));
  if (!out)
    return BSEG_E_ARG;
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  float* slot[EV_COUNT] = {nullptr};
  slot[EV_H2D] = &c->tm.h2d;
  slot[EV_BBOX] = &c->tm.bbox_keys;
  slot[EV_SORT] = &c->tm.sort;
  slot[EV_CELLS] = &c->tm.cells;
  slot[EV_KNN] = &c->tm.knn;
  slot[EV_KNN_FB] = &c->tm.knn_fallback;
  slot[EV_GROW] = &c->tm.grow;
  slot[EV_FINALIZE] = &c->tm.finalize;
  slot[EV_RASTER] = &c->tm.raster;
  slot[EV_D2H] = &c->tm.d2h;
  float total = 0.f;
  for (int i = 0; i < EV_COUNT; ++i) {
    if (!slot[i])
      continue;
    float ms = 0.f;
    if (c->ev_set[i])
      cudaEventElapsedTime(&ms, c->ev[i], c->ev_end[i]);
    *slot[i] = ms;
    total += ms;
  }
  c->tm.normals = 0.f;  // fused into the kNN kernel's epilogue
  c->tm.raster_host = c->tm_raster_host_ms;
  c->tm.total = total;
  c->tm.kernel_launches = c->launches;
  *out = c->tm;
  return 0;
}

BSEG_API int bseg_reset_counters(bseg_ctx* c)
{
  RC_CHECK(check_ctx(c));
  c->launches = 0;
  for (int i = 0; i < EV_COUNT; ++i)
    c->ev_set[i] = false;
  memset(&c->tm, 0, sizeof(c->tm));
  return 0;
}

BSEG_API void* bseg_stream(bseg_ctx* c) { return c ? (void*)c->stream : nullptr; }

BSEG_API int64_t bseg_point_count(const bseg_ctx* c) { return c ? c->n : -1; }

BSEG_API int32_t bseg_plane_count(const bseg_ctx* c) { return (c && c->have_grow) ? c->n_planes : -1; }

BSEG_API int bseg_debug_sort_pairs(bseg_ctx* c, uint64_t* keys, uint32_t* vals, int64_t n, int key_bits)
{
  RC_CHECK(check_ctx(c));
  if (n < 0 || key_bits < 0 || key_bits > 64 || (n > 0 && (!keys || !vals)))
    return bseg_fail(c, BSEG_E_ARG, "bseg_debug_sort_pairs: bad arguments");
  if (n == 0)
    return 0;
  for (int i = 0; i < 2; ++i) {
    RC_CHECK(dev_ensure(c, c->keys[i], (size_t)n * 8));
    RC_CHECK(dev_ensure(c, c->vals[i], (size_t)n * 4));
  }
  CU_CHECK(c, cudaMemcpyAsync(c->keys[0].p, keys, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  CU_CHECK(c, cudaMemcpyAsync(c->vals[0].p, vals, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  int sel = 0;
  RC_CHECK(bseg_sort_pairs_u64(c, dptr<uint64_t>(c->keys[0]), dptr<uint64_t>(c->keys[1]), dptr<uint32_t>(c->vals[0]),
                               dptr<uint32_t>(c->vals[1]), n, key_bits, &sel));
  CU_CHECK(c, cudaMemcpyAsync(keys, c->keys[sel].p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaMemcpyAsync(vals, c->vals[sel].p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  c->have_knn = c->have_grow = c->have_bin = false;
  return 0;
}

BSEG_API int bseg_debug_exclusive_scan(bseg_ctx* c, uint32_t* data, int64_t n)
{
  RC_CHECK(check_ctx(c));
  if (n < 0 || (n > 0 && !data))
    return bseg_fail(c, BSEG_E_ARG, "bseg_debug_exclusive_scan: bad arguments");
  if (n == 0)
    return 0;
  RC_CHECK(dev_ensure(c, c->flags, (size_t)(n + 4) * 4));
  CU_CHECK(c, cudaMemcpyAsync(c->flags.p, data, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  RC_CHECK(bseg_exclusive_scan_u32(c, dptr<uint32_t>(c->flags), n, nullptr));
  CU_CHECK(c, cudaMemcpyAsync(data, c->flags.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

}  // extern "C"
