// common.cuh -- context, device buffers and launch helpers shared by the sm_100a kernels.
#pragma once
#include <vector>

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bseg.h"

#define BSEG_NUM_SMS_FALLBACK 148
#define BSEG_MAX_K 32          // neighbours per row (lane-resident lists in the grower / fallback)
#define BSEG_MAX_COORD (1 << 23)  // shifted coordinates must stay below this: moment sums exact in fp64
#define BSEG_MAX_CELL 2048     // kNN cell edge limit: per-lane int32 moment partials cannot overflow

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

enum StageEv {
  EV_START = 0,
  EV_H2D,
  EV_BBOX,
  EV_SORT,
  EV_CELLS,
  EV_KNN,
  EV_KNN_FB,
  EV_GROW,
  EV_FINALIZE,
  EV_RASTER,
  EV_D2H,
  EV_COUNT
};

// per-plane record produced by the grower (device and host view)
struct PlaneRec {
  int32_t seed;       // original index of the seed
  int32_t pad;
  int64_t off;        // first entry in the list pool
  int64_t len;        // entries (duplicates kept)
  double nrm[3];      // cur_normal after the last Broad
  int32_t ctr[3];     // cur_center after the last Broad
  int32_t pad2;
};

struct bseg_ctx {
  int device = 0;
  int num_sms = BSEG_NUM_SMS_FALLBACK;
  cudaStream_t stream = nullptr;
  char err[512] = {0};

  // ---- cloud ----
  int64_t n = 0;        // points in the cloud (own + halo)
  int64_t n_owned = 0;  // points this rank owns (== n on one GPU)
  int32_t mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  bool have_points = false, have_knn = false, have_grow = false;
  bool normals_nonunit = false;  // caller-supplied normals that are not unit length (bseg_override_neigh_normals)
  bool have_bin = false;     // sorted cloud / cell tables valid for the current points and parameters
  bool have_origin = false;  // bseg_set_origin: shift by origin[] instead of the cloud's own minimum
  int32_t origin[3] = {0, 0, 0};
  int32_t grow_off[3] = {0, 0, 0};  // bseg_set_grow_offset: added to the shifted coordinates the grower sees

  // ---- binning state (valid after the knn stage) ----
  int32_t cell = 0;     // kNN cell edge
  int key_bits = 0;     // significant bits of the Morton key
  int64_t n_cells = 0;  // occupied cells
  uint64_t hash_mask = 0;
  int sort_sel = 0;     // which of keys[]/vals[] holds the sorted result
  int K = 0;
  // second, finer cell table (edge cell / 2) over the same sorted cloud: dense groups of the kNN descend to it
  bool have_mid = false;
  int64_t n_cells2 = 0;
  uint64_t hash_mask2 = 0;

  // ---- grow results ----
  int32_t n_planes = 0;
  int64_t n_plane_entries = 0;
  int64_t pool_cap = 0;
  void* h_planes = nullptr;  // std::vector<PlaneRec>*, host copy sorted by seed

  // ---- raster ----
  int32_t rW = 0, rH = 0;

  // ---- device buffers ----
  DevBuf xyz_raw;     // int32 [n][3], shifted to min=0, original order
  DevBuf minmax;      // int32 [8]
  DevBuf keys[2];     // u64 [n]
  DevBuf vals[2];     // u32 [n]
  DevBuf sort_cnt;    // u32 radix counters
  DevBuf scan_tmp;    // u32 block sums for the scan
  DevBuf pts;         // int4 [n] sorted: x,y,z,orig
  DevBuf inv;         // u32 [n]  orig -> sorted position
  DevBuf flags;       // u32 [n+1] scratch flags / scans
  DevBuf cell_key;    // u64 [n_cells]
  DevBuf cell_start;  // u32 [n_cells+1]
  DevBuf hash_keys;   // u64 [hash]
  DevBuf hash_vals;   // u32 [hash]
  DevBuf cell_key2, cell_start2, hash_keys2, hash_vals2;  // the finer table (have_mid)
  DevBuf nbr;         // int32 [n][K] sorted-position space, -1 padded
  DevBuf nrm;         // double [n][3] sorted-position space
  DevBuf curv;        // double [n]
  DevBuf worklist;    // u32 lists: overflow cells, unresolved queries
  DevBuf counters;    // u64 [64] misc counters
  DevBuf out_tmp;     // staging for original-order exports
  DevBuf x_neigh, x_normals;  // rows / normals in original order kept on the device (bseg_knn_device_results)
  // grower (sorted-position space unless noted)
  DevBuf g_state;     // int32 [n]: -1 free, else original index of the owning seed
  DevBuf g_res;       // (unused: the reservations live next to the owners in g_state, 8 bytes per point)
  DevBuf g_spec;      // u32 [n] (original index space): speculation record of the seed
  DevBuf g_pool;      // int32 list pool (sorted positions): pointIdx of running / committed planes
  DevBuf g_planes;    // PlaneRec []
  DevBuf g_tx;        // grower transaction slots
  DevBuf g_queue;     // u32 grower queue
  DevBuf g_stack;     // int2 DFS frames of grower slots
  DevBuf g_rowdup;    // u8 [n]: neighbour row names a point twice
  DevBuf g_marklog;   // uint2 [n]: orphan marks of the current sweep (point, seed)
  DevBuf g_label;     // int32 [n] label in original order
  DevBuf g_pidx;      // int32 [n] planeIdx in original order
  DevBuf g_pts_raw;   // int4 [n]: pts + grow_off (only when an offset is set)
  DevBuf g_nbr_masked;  // int32 [n][K]: rows cut at grow_radius (only when the radius-search variant is on)
  // raster
  DevBuf r_hist;
  DevBuf r_image;     // double [W*H*3]
  DevBuf r_png;       // u8 [3][W*H*3]
  DevBuf r_pix;       // u32 [W*H+1] pixel starts
  DevBuf r_cnt;       // double [W*H] weight sums of the count channel (finalised on the host)
  double* h_cnt = nullptr;  // pinned host copy of r_cnt, then log(count + 1) + bias
  size_t h_cnt_cap = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t raster_done = nullptr, raster_copied = nullptr;
  void* raster_worker = nullptr;  // std::thread* of an asynchronous host half
  double raster_max1 = 0.0;       // maximum of channel 1 after the host half
  float tm_raster_host_ms = 0.f;

  // ---- timing ----
  cudaEvent_t ev[EV_COUNT] = {nullptr};      // stage begin
  cudaEvent_t ev_end[EV_COUNT] = {nullptr};  // stage end
  bool ev_set[EV_COUNT] = {false};
  void* pinned = nullptr;  // small pinned host scratch for readbacks
  size_t pinned_cap = 0;
  bseg_timings tm;
  int64_t launches = 0;
  // per-device one-time setup (cudaFuncSetAttribute is per device, a context is bound to one)
  bool attr_sweep_set = false, attr_knn_set = false;
  cudaEvent_t grow_ev[12] = {nullptr};  // two sets of phase events of the speculative grower (created on first use)
  unsigned long long* grow_pinned = nullptr;  // pinned copies of the grower's control block, one per round in flight
  // the sweeper (high priority) and the background slice of the growers run side by side (grow_spec.cu)
  cudaStream_t out_stream = nullptr;  // bseg_segment_host: the shifted cloud travels back beside the kNN
  cudaEvent_t out_ready = nullptr;
  cudaStream_t grow_hi = nullptr, grow_lo = nullptr, grow_pf = nullptr;
  cudaEvent_t grow_fork = nullptr, grow_join[3] = {nullptr, nullptr, nullptr};
};

int bseg_fail(bseg_ctx* c, int code, const char* fmt, ...);

#define CU_CHECK(c, call)                                                                     \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return bseg_fail((c), BSEG_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,         \
                       cudaGetErrorString(e__));                                              \
  } while (0)

#define RC_CHECK(call)       \
  do {                       \
    int rc__ = (call);       \
    if (rc__ != 0)           \
      return rc__;           \
  } while (0)

#define KLAUNCH_CHECK(c)                                  \
  do {                                                    \
    (c)->launches++;                                      \
    CU_CHECK((c), cudaGetLastError());                    \
  } while (0)

int dev_ensure(bseg_ctx* c, DevBuf& b, size_t bytes);
void dev_free(DevBuf& b);

template <typename T>
static inline T* dptr(DevBuf& b)
{
  return reinterpret_cast<T*>(b.p);
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- primitives implemented in scan.cu / sort.cu ------------------------------------------------
// in-place exclusive prefix sum of u32; total (sum of all) written to *d_total when non-null
int bseg_exclusive_scan_u32(bseg_ctx* c, uint32_t* d_data, int64_t n, uint32_t* d_total);
// stable LSD radix sort of (key,val) pairs on key bits [0,key_bits); result left in k{out}/v{out}
int bseg_sort_pairs_u64(bseg_ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, int64_t n,
                        int key_bits, int* out_sel);
int bseg_sort_pairs_u32(bseg_ctx* c, uint32_t* k0, uint32_t* k1, uint32_t* v0, uint32_t* v1, int64_t n,
                        int key_bits, int* out_sel);

// ---- stages ---------------------------------------------------------------------------------------
int stage_bbox_shift(bseg_ctx* c);                           // bin.cu
int stage_bin(bseg_ctx* c, const bseg_params* p);            // bin.cu
int stage_knn(bseg_ctx* c, const bseg_params* p);            // knn.cu
int stage_export_knn(bseg_ctx* c, const bseg_params* p, int32_t* h_neigh, double* h_normals, double* h_curv);
int stage_halo_check(bseg_ctx* c, int32_t x_lo, int32_t x_hi, int32_t halo, int64_t* n_unresolved);  // knn.cu
int stage_override(bseg_ctx* c, const bseg_params* p, const int32_t* neigh, const double* normals, bool on_device);
int stage_export_knn_device(bseg_ctx* c, const bseg_params* p, const int32_t** d_neigh, const double** d_normals);
int stage_alloc_knn_outputs(bseg_ctx* c, const bseg_params* p);
int stage_grow(bseg_ctx* c, const bseg_params* p);           // grow.cu
void grow_host_free(bseg_ctx* c);
int stage_export_grow(bseg_ctx* c, int32_t* h_plane_idx, int32_t* h_label);
int stage_get_planes(bseg_ctx* c, int32_t* seeds, double* normals, int32_t* centers, int64_t* offsets,
                     int32_t* point_idx);
int stage_paint(bseg_ctx* c, const int32_t* h_ids, int32_t n_listed, const uint16_t* h_rgb, uint16_t* h_colors);
int stage_plane_classes(bseg_ctx* c, double facade_max_nz, double roof_min_nz, double ground_z, double* h_eq,
                        uint8_t* h_plane_class, uint8_t* h_point_class);
int stage_contour_mask(bseg_ctx* c, const uint8_t* h_pixels, int32_t W, int32_t H, int32_t comp, int32_t channel, int32_t thresh,
                       int32_t iterations, uint8_t* h_mask);  // contour.cu
void contour_find_host(const uint8_t* mask, int32_t W, int32_t H, bool simple, std::vector<int32_t>& pts, std::vector<int64_t>& offsets);
double contour_area(const int32_t* xy, int64_t n);
double contour_perimeter(const int32_t* xy, int64_t n);
void contour_draw(uint8_t* img, int32_t W, int32_t H, int32_t comp, const int32_t* xy, int64_t n, const uint8_t* color);
int stage_raster_size(bseg_ctx* c, const bseg_params* p, int32_t* W, int32_t* H);
// mode: host part of the count channel done before returning (SYNC), on a worker thread that raster_host_join()
// waits for (ASYNC: overlaps the grower), or left to the caller (DEVICE_ONLY: bseg_raster_device)
enum { RASTER_SYNC = 0, RASTER_ASYNC = 1, RASTER_DEVICE_ONLY = 2 };
int stage_raster(bseg_ctx* c, const bseg_params* p, double* h_image, uint8_t* a, uint8_t* b, uint8_t* cc,
                 double* th, int mode, const double* th_override = nullptr);
int raster_host_join(bseg_ctx* c);
double bseg_count_channel_host(double* v, int64_t n, double bias);
int stage_label_raster(bseg_ctx* c, const bseg_params* p, const uint16_t* h_plane_rgb, int32_t* h_label, uint8_t* h_rgb);

#define STAGE_BEGIN(c, which) cudaEventRecord((c)->ev[which], (c)->stream)
#define STAGE_END(c, which)                            \
  do {                                                 \
    cudaEventRecord((c)->ev_end[which], (c)->stream);  \
    (c)->ev_set[which] = true;                         \
  } while (0)

// small device -> host readback through pinned memory (synchronises the stream)
int read_back(bseg_ctx* c, void* host_dst, const void* dev_src, size_t bytes);

// ---- device helpers ---------------------------------------------------------------------------------
#ifdef __CUDACC__
#define FULL_MASK 0xffffffffu

__device__ __forceinline__ uint64_t morton_spread21(uint32_t v)
{
  uint64_t x = v & 0x1fffffULL;
  x = (x | x << 32) & 0x1f00000000ffffULL;
  x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL;
  x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}
__device__ __forceinline__ uint32_t morton_compact21(uint64_t x)
{
  x &= 0x1249249249249249ULL;
  x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ULL;
  x = (x ^ (x >> 4)) & 0x100f00f00f00f00fULL;
  x = (x ^ (x >> 8)) & 0x1f0000ff0000ffULL;
  x = (x ^ (x >> 16)) & 0x1f00000000ffffULL;
  x = (x ^ (x >> 32)) & 0x1fffffULL;
  return (uint32_t)x;
}
__device__ __forceinline__ uint64_t morton3(uint32_t x, uint32_t y, uint32_t z)
{
  return morton_spread21(x) | (morton_spread21(y) << 1) | (morton_spread21(z) << 2);
}
__device__ __forceinline__ uint64_t hash64(uint64_t k)
{
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}
#define HASH_EMPTY 0xffffffffffffffffULL
__device__ __forceinline__ uint32_t hash_lookup(const uint64_t* __restrict__ hk, const uint32_t* __restrict__ hv,
                                                uint64_t mask, uint64_t key)
{
  uint64_t h = hash64(key) & mask;
  for (;;) {
    uint64_t k = __ldg(hk + h);
    if (k == key)
      return __ldg(hv + h);
    if (k == HASH_EMPTY)
      return 0xffffffffu;
    h = (h + 1) & mask;
  }
}
__device__ __forceinline__ uint32_t lanemask_lt()
{
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
#endif
