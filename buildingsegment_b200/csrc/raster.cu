// raster.cu -- north_star stage (5): ground threshold, height / count rasterisation, PNG pixels.
//
// Replaces buildingSeg::groundTH (TMC3.cpp:181-198), compute_gird_picture (:127-172) and the pixel
// part of save_image (:81-121); the PNG encode itself stays with stb_image_write on the host.
//
// The reference accumulates fp64 sums per pixel IN POINT ORDER (:132-148), so atomics would change
// the rounding.  Instead: kept points (z >= th) are stably radix-sorted by their source pixel
// (y/bin * W + x/bin), which leaves every pixel's points in original index order; one thread per
// OUTPUT pixel then merges the four source pixels that splat into it (taps (0,0) (1,0) (0,1) (1,1))
// by ascending point index and adds s and s*z exactly as the reference does.  Bit-exact, no atomics
// on doubles.  Then mean height, per-channel max and the uint8 images.
//
// The count channel is finalised on the HOST: log(count + 1) is std::log of the platform's libm in the reference
// (TMC3.cpp:161), no device polynomial reproduces another libm's last bit, and a 1-ulp move can flip
// (uint8)(255.0 * v / max).  The device produces the exact weight sums; raster_count_channel_host() takes them
// through the same libm the reference links, on worker threads, and -- in the whole-path calls -- while the plane
// grower runs on the GPU.  The doubles and bytes are then the reference's, bit for bit.
//
// Traffic: histogram R 4 B/pt; keys R 12 W 8 B/pt; sort 4 passes x 16 B/pt; accumulate gathers
// 12 B per tap (4 taps/pt) and writes 24 B/pixel; PNG pass R 24 W 9 B/pixel.
#include <cmath>
#include <chrono>
#include <thread>
#include <vector>

#include "common.cuh"
#include "bseg_arith.h"

namespace {

constexpr int TPB = 256;

__global__ void __launch_bounds__(TPB) zhist_kernel(const int32_t* __restrict__ xyz, int64_t n, int32_t bin_height,
                                                    uint32_t* __restrict__ hist, int nb_bins)
{
  extern __shared__ uint32_t sh[];
  const bool use_sh = nb_bins <= 4096;
  if (use_sh) {
    for (int i = threadIdx.x; i < nb_bins; i += TPB) sh[i] = 0;
    __syncthreads();
  }
  int64_t stride = (int64_t)gridDim.x * TPB;
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) {
    int b = xyz[3 * i + 2] / bin_height;
    if (use_sh) atomicAdd(&sh[b], 1u);
    else atomicAdd(&hist[b], 1u);
  }
  if (use_sh) {
    __syncthreads();
    for (int i = threadIdx.x; i < nb_bins; i += TPB)
      if (sh[i]) atomicAdd(&hist[i], sh[i]);
  }
}

__global__ void __launch_bounds__(TPB) pixkey_kernel(const int32_t* __restrict__ xyz, int64_t n, int32_t bin, int32_t W,
                                                     int32_t th, uint32_t sentinel, uint32_t* __restrict__ keys,
                                                     uint32_t* __restrict__ vals)
{
  int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n)
    return;
  int32_t x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  keys[i] = (z < th) ? sentinel : (uint32_t)((y / bin) * W + (x / bin));
  vals[i] = (uint32_t)i;
}

// start[px] = first sorted position whose key >= px, for px in 0..P (P = sentinel)
__global__ void __launch_bounds__(TPB) pixstart_kernel(const uint32_t* __restrict__ keys, int64_t n, uint32_t P,
                                                       uint32_t* __restrict__ start)
{
  int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (t > n)
    return;
  uint32_t k = t < n ? min(keys[t], P) : P;
  int64_t kp = t > 0 ? (int64_t)min(keys[t - 1], P) : -1;
  for (int64_t j = kp + 1; j <= (int64_t)k; ++j)
    start[j] = (uint32_t)t;
}

struct Run {
  uint32_t cur, end;
  int xi, yi;
};

__global__ void __launch_bounds__(TPB) accumulate_kernel(const int32_t* __restrict__ xyz, const uint32_t* __restrict__ vals,
                                                         const uint32_t* __restrict__ start, int32_t W, int32_t H,
                                                         int32_t bin, double* __restrict__ image,
                                                         double* __restrict__ counts,
                                                         unsigned long long* __restrict__ maxbits)
{
  int64_t px = (int64_t)blockIdx.x * TPB + threadIdx.x;
  double c0 = 0.0, c1 = 0.0;
  if (px < (int64_t)W * H) {
    const int X = (int)(px % W), Y = (int)(px / W);
    Run r[4];
    int nr = 0;
    for (int yi = 0; yi < 2; ++yi)
      for (int xi = 0; xi < 2; ++xi) {
        int sx = X - xi, sy = Y - yi;
        if (sx >= 0 && sy >= 0) {
          uint32_t sp = (uint32_t)(sy * W + sx);
          uint32_t a = start[sp], b = start[sp + 1];
          if (a < b) {
            r[nr].cur = a; r[nr].end = b; r[nr].xi = xi; r[nr].yi = yi;
            ++nr;
          }
        }
      }
    uint32_t head[4];
    for (int k = 0; k < 4; ++k)
      head[k] = k < nr ? vals[r[k].cur] : 0xffffffffu;
    for (;;) {
      int best = -1;
      uint32_t bi = 0xffffffffu;
      for (int k = 0; k < 4; ++k)
        if (k < nr && head[k] < bi) { bi = head[k]; best = k; }
      if (best < 0)
        break;
      const int32_t p0 = xyz[3 * (int64_t)bi], p1 = xyz[3 * (int64_t)bi + 1], p2 = xyz[3 * (int64_t)bi + 2];
      const int x = p0 / bin, y = p1 / bin;
      const double w = 1.0 * p0 / bin - x;  // TMC3.cpp:141-142
      const double h = 1.0 * p1 / bin - y;
      const double s = ((r[best].xi == 1) ? w : (1 - w)) * ((r[best].yi == 1) ? h : (1 - h));  // :143
      c1 += s;        // :144
      c0 += s * p2;   // :145
      ++r[best].cur;
      head[best] = r[best].cur < r[best].end ? vals[r[best].cur] : 0xffffffffu;
    }
    if (c1 != 0) c0 = c0 / c1;          // :152-157
    image[3 * px] = c0;
    image[3 * px + 1] = c1;             // the exact weight sum: log(c1 + 1) (+ bias) is taken on the host (:161-163)
    image[3 * px + 2] = 0.0;
    counts[px] = c1;
  }
  // per-channel maximum (values are >= 0, so the bit patterns order like the doubles)
  unsigned long long b0 = (unsigned long long)__double_as_longlong(c0 > 0 ? c0 : 0.0);
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t0 = __shfl_xor_sync(FULL_MASK, b0, o);
    b0 = t0 > b0 ? t0 : b0;
  }
  if ((threadIdx.x & 31) == 0) {
    if (b0) atomicMax(&maxbits[0], b0);
  }
}

// TMC3.cpp:93-98,112-119: images A (mean height -> byte 0) and C (always zero); image B comes from the host
__global__ void __launch_bounds__(TPB) png_kernel(const double* __restrict__ image, int64_t npx,
                                                  const unsigned long long* __restrict__ maxbits,
                                                  uint8_t* __restrict__ a, uint8_t* __restrict__ b,
                                                  uint8_t* __restrict__ cimg)
{
  int64_t px = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (px >= npx)
    return;
  const double m0 = __longlong_as_double((long long)maxbits[0]);
  uint8_t va = 0;
  if (m0 != 0) va = (uint8_t)(255.0 * (1.0 * image[3 * px] / m0));
  a[3 * px] = va; a[3 * px + 1] = 0; a[3 * px + 2] = 0;
  b[3 * px] = 0; b[3 * px + 1] = 0; b[3 * px + 2] = 0;
  cimg[3 * px] = 0; cimg[3 * px + 1] = 0; cimg[3 * px + 2] = 0;
}

// ---- host side of the count channel ------------------------------------------------------------------------------
template <class F>
void parallel_ranges(int64_t n, F f)
{
  unsigned hw = std::thread::hardware_concurrency();
  int nt = hw ? (int)hw : 4;
  if (nt > 32) nt = 32;
  if (n < 65536) nt = 1;
  if (nt <= 1) {
    f(0, (int64_t)0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t per = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int64_t lo = t * per, hi = lo + per < n ? lo + per : n;
    if (lo >= hi)
      break;
    th.emplace_back(f, t, lo, hi);
  }
  for (auto& x : th) x.join();
}

}  // namespace

// v[i] = log(v[i] + 1); if (v[i] != 0) v[i] += bias   (TMC3.cpp:159-164) with the platform's std::log; returns the maximum
double bseg_count_channel_host(double* v, int64_t n, double bias)
{
  double part[32];
  for (double& x : part) x = 0.0;
  parallel_ranges(n, [&](int t, int64_t lo, int64_t hi) {
    double m = 0.0;
    for (int64_t i = lo; i < hi; ++i) {
      double x = v[i];
      if (x != 0.0) {  // log(0 + 1) == 0 exactly and stays 0 (:162)
        x = std::log(x + 1);
        if (x != 0) x += bias;
        v[i] = x;
      }
      if (m < x) m = x;
    }
    part[t] = m;
  });
  double m = 0.0;
  for (double x : part)
    if (m < x) m = x;
  return m;
}

namespace {

// everything after the device pass, on the calling thread or a worker: wait for the sums, libm log, maximum,
// image B bytes (TMC3.cpp:101-108) and channel 1 of the caller's double image
void raster_host_job(bseg_ctx* c, int64_t npx, double bias, double* h_image, uint8_t* h_b)
{
  const auto t0 = std::chrono::steady_clock::now();
  cudaSetDevice(c->device);
  cudaEventSynchronize(c->raster_copied);
  double* v = c->h_cnt;
  const double mx = bseg_count_channel_host(v, npx, bias);
  c->raster_max1 = mx;
  if (h_image || h_b)
    parallel_ranges(npx, [&](int, int64_t lo, int64_t hi) {
      for (int64_t px = lo; px < hi; ++px) {
        if (h_image) h_image[3 * px + 1] = v[px];
        if (h_b) {
          h_b[3 * px] = 0;
          h_b[3 * px + 1] = mx != 0 ? (uint8_t)(255.0 * (1.0 * v[px] / mx)) : (uint8_t)0;
          h_b[3 * px + 2] = 0;
        }
      }
    });
  c->tm_raster_host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

int stage_raster_size(bseg_ctx* c, const bseg_params* p, int32_t* W, int32_t* H)
{
  int32_t w = 0, h = 0;
  if (c->n > 0) {
    w = (c->mx[0] - c->mn[0]) / p->bin + 2;  // TMC3.cpp:75-76
    h = (c->mx[1] - c->mn[1]) / p->bin + 2;
  }
  c->rW = w;
  c->rH = h;
  if (W) *W = w;
  if (H) *H = h;
  return 0;
}

int stage_raster(bseg_ctx* c, const bseg_params* p, double* h_image, uint8_t* h_a, uint8_t* h_b, uint8_t* h_c,
                 double* h_th, int mode, const double* th_override)
{
  const int64_t n = c->n;
  RC_CHECK(stage_raster_size(c, p, nullptr, nullptr));
  if (n == 0) {
    if (h_th) *h_th = 0.0;
    return 0;
  }
  const int32_t W = c->rW, H = c->rH;
  const int64_t npx = (int64_t)W * H;
  if (npx >= ((int64_t)1 << 31) - 2)
    return bseg_fail(c, BSEG_E_ARG, "raster of %d x %d pixels exceeds 2^31", W, H);
  const int32_t zext = c->mx[2] - c->mn[2];
  const int nb_bins = zext / p->bin_height + 1;

  STAGE_BEGIN(c, EV_RASTER);
  // ---- groundTH: z histogram, first bin where the cumulative count exceeds N/2 (or the caller's threshold:
  //      the slabs of a tile share the tile's, bseg_raster_device) ----
  int32_t th;
  if (th_override) {
    th = (int32_t)*th_override;
  } else {
    RC_CHECK(dev_ensure(c, c->r_hist, (size_t)nb_bins * 4 + 64));
    uint32_t* d_hist = dptr<uint32_t>(c->r_hist);
    CU_CHECK(c, cudaMemsetAsync(d_hist, 0, (size_t)nb_bins * 4, c->stream));
    {
      int g = (int)ceil_div64(n, TPB * 8);
      if (g > c->num_sms * 8) g = c->num_sms * 8;
      if (g < 1) g = 1;
      size_t sh = nb_bins <= 4096 ? (size_t)nb_bins * 4 : 0;
      zhist_kernel<<<g, TPB, sh, c->stream>>>(dptr<int32_t>(c->xyz_raw), n, p->bin_height, d_hist, nb_bins);
      KLAUNCH_CHECK(c);
    }
    std::vector<uint32_t> hist((size_t)nb_bins);
    RC_CHECK(read_back(c, hist.data(), d_hist, (size_t)nb_bins * 4));
    const int TH = (int)(n / 2);
    int total = 0;
    int i;
    for (i = 0; i < nb_bins; ++i) {
      total += (int)hist[(size_t)i];
      if (total > TH)
        break;
    }
    th = i * p->bin_height;
  }
  if (h_th) *h_th = (double)th;

  // ---- source-pixel keys, stable sort, pixel starts ----
  for (int i = 0; i < 2; ++i) {
    RC_CHECK(dev_ensure(c, c->keys[i], (size_t)n * 8));
    RC_CHECK(dev_ensure(c, c->vals[i], (size_t)n * 4));
  }
  // keys[] doubles as u32 key storage here; this invalidates the kNN stage's sorted Morton keys,
  // which only the (already finished) fallback needed
  uint32_t* k0 = dptr<uint32_t>(c->keys[0]);
  uint32_t* k1 = dptr<uint32_t>(c->keys[1]);
  uint32_t* v0 = dptr<uint32_t>(c->vals[0]);
  uint32_t* v1 = dptr<uint32_t>(c->vals[1]);
  const unsigned nbk = (unsigned)ceil_div64(n, TPB);
  pixkey_kernel<<<nbk, TPB, 0, c->stream>>>(dptr<int32_t>(c->xyz_raw), n, p->bin, W, th, (uint32_t)npx, k0, v0);
  KLAUNCH_CHECK(c);
  int bits = 1;
  while (((int64_t)1 << bits) <= npx) ++bits;
  int sel = 0;
  RC_CHECK(bseg_sort_pairs_u32(c, k0, k1, v0, v1, n, bits, &sel));
  const uint32_t* keys = sel ? k1 : k0;
  const uint32_t* vals = sel ? v1 : v0;
  RC_CHECK(dev_ensure(c, c->r_pix, (size_t)(npx + 2) * 4));
  pixstart_kernel<<<(unsigned)ceil_div64(n + 1, TPB), TPB, 0, c->stream>>>(keys, n, (uint32_t)npx,
                                                                         dptr<uint32_t>(c->r_pix));
  KLAUNCH_CHECK(c);

  // ---- ordered accumulation + finalisation ----
  RC_CHECK(dev_ensure(c, c->r_image, (size_t)npx * 24));
  RC_CHECK(dev_ensure(c, c->r_png, (size_t)npx * 9 + 64));
  RC_CHECK(dev_ensure(c, c->counters, 192 * sizeof(uint64_t)));
  unsigned long long* maxbits = reinterpret_cast<unsigned long long*>(dptr<uint64_t>(c->counters)) + 56;
  CU_CHECK(c, cudaMemsetAsync(maxbits, 0, 2 * sizeof(unsigned long long), c->stream));
  RC_CHECK(dev_ensure(c, c->r_cnt, (size_t)npx * 8));
  accumulate_kernel<<<(unsigned)ceil_div64(npx, TPB), TPB, 0, c->stream>>>(
      dptr<int32_t>(c->xyz_raw), vals, dptr<uint32_t>(c->r_pix), W, H, p->bin, dptr<double>(c->r_image),
      dptr<double>(c->r_cnt), maxbits);
  KLAUNCH_CHECK(c);
  uint8_t* da = dptr<uint8_t>(c->r_png);
  uint8_t* db = da + 3 * npx;
  uint8_t* dc = db + 3 * npx;
  png_kernel<<<(unsigned)ceil_div64(npx, TPB), TPB, 0, c->stream>>>(dptr<double>(c->r_image), npx, maxbits, da, db, dc);
  KLAUNCH_CHECK(c);
  STAGE_END(c, EV_RASTER);

  // ---- the count channel: sums -> pinned host memory on the copy stream, then libm log on the host ----
  RC_CHECK(raster_host_join(c));  // a previous job still owns h_cnt
  if ((size_t)npx * 8 > c->h_cnt_cap) {
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    c->h_cnt = nullptr;
    c->h_cnt_cap = 0;
    const size_t want = (size_t)npx * 8 + (size_t)npx / 2 + 4096;
    CU_CHECK(c, cudaMallocHost((void**)&c->h_cnt, want));
    c->h_cnt_cap = want;
  }
  if (!c->copy_stream) CU_CHECK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  if (!c->raster_copied) CU_CHECK(c, cudaEventCreateWithFlags(&c->raster_copied, cudaEventDisableTiming));
  if (!c->raster_done) CU_CHECK(c, cudaEventCreateWithFlags(&c->raster_done, cudaEventDisableTiming));
  CU_CHECK(c, cudaEventRecord(c->raster_done, c->stream));
  CU_CHECK(c, cudaStreamWaitEvent(c->copy_stream, c->raster_done, 0));
  CU_CHECK(c, cudaMemcpyAsync(c->h_cnt, c->r_cnt.p, (size_t)npx * 8, cudaMemcpyDeviceToHost, c->copy_stream));
  if (h_image) CU_CHECK(c, cudaMemcpyAsync(h_image, c->r_image.p, (size_t)npx * 24, cudaMemcpyDeviceToHost, c->copy_stream));
  if (h_a) CU_CHECK(c, cudaMemcpyAsync(h_a, da, (size_t)npx * 3, cudaMemcpyDeviceToHost, c->copy_stream));
  if (h_c) CU_CHECK(c, cudaMemcpyAsync(h_c, dc, (size_t)npx * 3, cudaMemcpyDeviceToHost, c->copy_stream));
  CU_CHECK(c, cudaEventRecord(c->raster_copied, c->copy_stream));
  if (mode == RASTER_DEVICE_ONLY) {
    // bseg_raster_device: the caller finishes the count channel itself (the slabs of a tile share one maximum)
    CU_CHECK(c, cudaEventSynchronize(c->raster_copied));
    return 0;
  }
  const double bias = p->count_bias;
  if (mode == RASTER_ASYNC) {
    c->raster_worker = new std::thread(raster_host_job, c, npx, bias, h_image, h_b);
    return 0;
  }
  raster_host_job(c, npx, bias, h_image, h_b);
  return 0;
}

// wait for the host half of an asynchronous raster (bseg_run_device / bseg_segment_host overlap it with the grower)
int raster_host_join(bseg_ctx* c)
{
  if (c->raster_worker) {
    std::thread* t = static_cast<std::thread*>(c->raster_worker);
    t->join();
    delete t;
    c->raster_worker = nullptr;
  }
  return 0;
}

// ---- label raster (north_star stage 5; the reference has no counterpart: its images are height / count) ------
// Pixel (x / bin, y / bin) of the W x H grid of the height map shows the plane label of its HIGHEST point
// (ties: the lower original index); 0 where the pixel is empty or that point belongs to no plane.  One
// atomicMax per point on a packed (z, ~index) key, then one pass over the pixels.
namespace {
__global__ void __launch_bounds__(TPB) label_top_kernel(const int32_t* __restrict__ xyz, int64_t n, int32_t bin, int32_t W,
                                                        unsigned long long* __restrict__ top)
{
  int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n)
    return;
  const int32_t x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  const unsigned long long key = ((unsigned long long)(uint32_t)z << 32) | (uint32_t)~(uint32_t)i;
  atomicMax(top + (int64_t)(y / bin) * W + (x / bin), key + 1ull);  // 0 = empty pixel
}

__global__ void __launch_bounds__(TPB) label_pixels_kernel(const unsigned long long* __restrict__ top, int64_t npx,
                                                           const int32_t* __restrict__ label,
                                                           const uint16_t* __restrict__ plane_rgb,
                                                           int32_t* __restrict__ out_label, uint8_t* __restrict__ out_rgb)
{
  int64_t px = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (px >= npx)
    return;
  const unsigned long long t = top[px];
  int32_t l = 0;
  if (t) l = label[~(uint32_t)((t - 1ull) & 0xffffffffull)];
  out_label[px] = l;
  if (out_rgb) {
    uint8_t r = 0, g = 0, b = 0;
    if (l > 0 && plane_rgb) {
      r = (uint8_t)plane_rgb[3 * (int64_t)(l - 1)];
      g = (uint8_t)plane_rgb[3 * (int64_t)(l - 1) + 1];
      b = (uint8_t)plane_rgb[3 * (int64_t)(l - 1) + 2];
    }
    out_rgb[3 * px] = r;
    out_rgb[3 * px + 1] = g;
    out_rgb[3 * px + 2] = b;
  }
}
}  // namespace

int stage_label_raster(bseg_ctx* c, const bseg_params* p, const uint16_t* h_plane_rgb, int32_t* h_label, uint8_t* h_rgb)
{
  int32_t W = 0, H = 0;
  RC_CHECK(stage_raster_size(c, p, &W, &H));
  const int64_t n = c->n, npx = (int64_t)W * H;
  if (n == 0 || npx == 0)
    return 0;
  const size_t rgb_bytes = (size_t)c->n_planes * 6;
  // top keys | labels | rgb | plane colours
  RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)npx * (8 + 4 + 3) + rgb_bytes + 64));
  unsigned long long* top = dptr<unsigned long long>(c->out_tmp);
  int32_t* d_label = reinterpret_cast<int32_t*>(top + npx);
  uint16_t* d_prgb = reinterpret_cast<uint16_t*>(d_label + npx);
  uint8_t* d_rgb = reinterpret_cast<uint8_t*>(d_prgb) + ((rgb_bytes + 15) & ~(size_t)15);
  CU_CHECK(c, cudaMemsetAsync(top, 0, (size_t)npx * 8, c->stream));
  if (h_plane_rgb && rgb_bytes)
    CU_CHECK(c, cudaMemcpyAsync(d_prgb, h_plane_rgb, rgb_bytes, cudaMemcpyHostToDevice, c->stream));
  label_top_kernel<<<(unsigned)ceil_div64(n, TPB), TPB, 0, c->stream>>>(dptr<int32_t>(c->xyz_raw), n, p->bin, W, top);
  KLAUNCH_CHECK(c);
  label_pixels_kernel<<<(unsigned)ceil_div64(npx, TPB), TPB, 0, c->stream>>>(
      top, npx, dptr<int32_t>(c->g_label), (h_plane_rgb && rgb_bytes) ? d_prgb : nullptr, d_label, h_rgb ? d_rgb : nullptr);
  KLAUNCH_CHECK(c);
  if (h_label) CU_CHECK(c, cudaMemcpyAsync(h_label, d_label, (size_t)npx * 4, cudaMemcpyDeviceToHost, c->stream));
  if (h_rgb) CU_CHECK(c, cudaMemcpyAsync(h_rgb, d_rgb, (size_t)npx * 3, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}
