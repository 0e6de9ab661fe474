// sort.cu -- hand-written stable LSD radix sort of (key, u32 value) pairs, 8-bit digits.
//
// Per pass: (1) per-block digit histogram, (2) exclusive scan of the digit-major counter matrix,
// (3) re-read the tile, rank every key with warp-level match_any + per-warp digit counters,
// stage the tile in shared memory in digit order, and write each digit's run out contiguously
// (coalesced).  Traffic per pass: read keys twice, values once, write both once.
//
// Replaces nothing in the reference (its spatial index is Open3D's KD-tree, my_function.h:63,71);
// it is stage (1) of BASELINE.json's north_star: Morton / voxel binning by radix sort.
#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_RADIX = 256;
// keys per thread: the scatter tile (keys + values + counters) must fit 48 KB of static shared memory
template <typename KeyT>
struct RsCfg {
  static constexpr int ITEMS = sizeof(KeyT) == 8 ? 12 : 16;
  static constexpr int TILE = RS_THREADS * ITEMS;
};

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, int shift)
{
  return (uint32_t)(k >> shift) & 0xffu;
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const KeyT* __restrict__ keys, int64_t n, int shift,
                                                            uint32_t* __restrict__ counts, uint32_t nblocks)
{
  __shared__ uint32_t h[RS_RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  constexpr int RS_ITEMS = RsCfg<KeyT>::ITEMS;
  constexpr int RS_TILE = RsCfg<KeyT>::TILE;
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < n)
      atomicAdd(&h[digit_of(keys[i], shift)], 1u);
  }
  __syncthreads();
  counts[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const KeyT* __restrict__ keys_in,
                                                               const uint32_t* __restrict__ vals_in,
                                                               KeyT* __restrict__ keys_out,
                                                               uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                               const uint32_t* __restrict__ offsets, uint32_t nblocks)
{
  constexpr int RS_ITEMS = RsCfg<KeyT>::ITEMS;
  constexpr int RS_TILE = RsCfg<KeyT>::TILE;
  __shared__ uint32_t warp_cnt[RS_WARPS][RS_RADIX];  // per-warp digit counters -> per-warp bases
  __shared__ uint32_t dig_off[RS_RADIX];             // exclusive offset of each digit inside the tile
  __shared__ uint32_t dig_glob[RS_RADIX];            // global output offset of each digit for this block
  __shared__ KeyT s_keys[RS_TILE];
  __shared__ uint32_t s_vals[RS_TILE];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  const int64_t wbase = base + (int64_t)w * (RS_ITEMS * 32);
  const int tile_n = (int)((n - base) < RS_TILE ? (n - base) : RS_TILE);

  for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS)
    (&warp_cnt[0][0])[i] = 0;
  dig_glob[threadIdx.x] = offsets[(size_t)threadIdx.x * nblocks + blockIdx.x];
  __syncthreads();

  KeyT k[RS_ITEMS];
  uint32_t v[RS_ITEMS];
  uint32_t rank[RS_ITEMS];
  // warp-striped: item r of lane l is element wbase + r*32 + l, so (r, l) order is index order
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t i = wbase + r * 32 + lane;
    bool ok = i < n;
    k[r] = ok ? keys_in[i] : (KeyT)~(KeyT)0;
    v[r] = ok ? vals_in[i] : 0u;
  }
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t i = wbase + r * 32 + lane;
    bool ok = i < n;
    uint32_t d = ok ? digit_of(k[r], shift) : 0x100u;  // out-of-range lanes form their own group
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t lt = peers & ((1u << lane) - 1u);
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (ok && lane == leader) {
      old = warp_cnt[w][d];
      warp_cnt[w][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(lt);
    __syncwarp();
  }
  __syncthreads();
  // thread d: exclusive prefix of digit d over the warps, and the tile total of digit d
  {
    uint32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) {
      uint32_t t = warp_cnt[ww][threadIdx.x];
      warp_cnt[ww][threadIdx.x] = run;
      run += t;
    }
    // block-wide exclusive scan of the 256 digit totals
    uint32_t inc = run;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __shared__ uint32_t wt[RS_WARPS];
    if (lane == 31) wt[w] = inc;
    __syncthreads();
    uint32_t wb = 0;
    for (int ww = 0; ww < w; ++ww) wb += wt[ww];
    dig_off[threadIdx.x] = wb + inc - run;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      uint32_t d = digit_of(k[r], shift);
      uint32_t pos = dig_off[d] + warp_cnt[w][d] + rank[r];
      s_keys[pos] = k[r];
      s_vals[pos] = v[r];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < tile_n; j += RS_THREADS) {
    KeyT kk = s_keys[j];
    uint32_t d = digit_of(kk, shift);
    size_t o = (size_t)dig_glob[d] + (uint32_t)(j - dig_off[d]);
    keys_out[o] = kk;
    vals_out[o] = s_vals[j];
  }
}

template <typename KeyT>
int sort_pairs(bseg_ctx* c, KeyT* k0, KeyT* k1, uint32_t* v0, uint32_t* v1, int64_t n, int key_bits, int* out_sel)
{
  *out_sel = 0;
  if (n <= 1 || key_bits <= 0)
    return 0;
  if (n >= ((int64_t)1 << 32))
    return bseg_fail(c, BSEG_E_ARG, "sort: n >= 2^32");
  int passes = (key_bits + 7) / 8;
  constexpr int RS_TILE = RsCfg<KeyT>::TILE;
  uint32_t nb = (uint32_t)ceil_div64(n, RS_TILE);
  size_t ncnt = (size_t)RS_RADIX * nb;
  RC_CHECK(dev_ensure(c, c->sort_cnt, (ncnt + 4) * sizeof(uint32_t)));
  uint32_t* cnt = dptr<uint32_t>(c->sort_cnt);
  KeyT* kin = k0;
  KeyT* kout = k1;
  uint32_t* vin = v0;
  uint32_t* vout = v1;
  for (int p = 0; p < passes; ++p) {
    int shift = p * 8;
    rs_hist_kernel<KeyT><<<nb, RS_THREADS, 0, c->stream>>>(kin, n, shift, cnt, nb);
    KLAUNCH_CHECK(c);
    RC_CHECK(bseg_exclusive_scan_u32(c, cnt, (int64_t)ncnt, nullptr));
    rs_scatter_kernel<KeyT><<<nb, RS_THREADS, 0, c->stream>>>(kin, vin, kout, vout, n, shift, cnt, nb);
    KLAUNCH_CHECK(c);
    KeyT* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
    *out_sel ^= 1;
  }
  return 0;
}

}  // namespace

int bseg_sort_pairs_u64(bseg_ctx* c, uint64_t* k0, uint64_t* k1, uint32_t* v0, uint32_t* v1, int64_t n,
                        int key_bits, int* out_sel)
{
  return sort_pairs<uint64_t>(c, k0, k1, v0, v1, n, key_bits, out_sel);
}

int bseg_sort_pairs_u32(bseg_ctx* c, uint32_t* k0, uint32_t* k1, uint32_t* v0, uint32_t* v1, int64_t n,
                        int key_bits, int* out_sel)
{
  return sort_pairs<uint32_t>(c, k0, k1, v0, v1, n, key_bits, out_sel);
}
