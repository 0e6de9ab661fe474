// scan.cu -- device-wide exclusive prefix sum (u32), reduce / recurse / downsweep.
// Used by the radix sort (digit offsets), the cell table (head flags -> cell ids) and the raster
// (kept-point compaction).  HBM-bound: 2 reads + 1 write of the array.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ROUNDS = 4;
constexpr int SCAN_CHUNK = SCAN_THREADS * 4 * SCAN_ROUNDS;  // 4096 elements per block

__device__ __forceinline__ uint4 load4(const uint32_t* __restrict__ d, int64_t i, int64_t n)
{
  uint4 v = make_uint4(0, 0, 0, 0);
  if (i + 3 < n) {
    v = *reinterpret_cast<const uint4*>(d + i);
  } else {
    if (i < n) v.x = d[i];
    if (i + 1 < n) v.y = d[i + 1];
    if (i + 2 < n) v.z = d[i + 2];
  }
  return v;
}

__device__ __forceinline__ uint32_t block_reduce(uint32_t v, uint32_t* sm)
{
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0)
    sm[w] = v;
  __syncthreads();
  uint32_t t = 0;
  for (int i = 0; i < SCAN_THREADS / 32; ++i)
    t += sm[i];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ d, int64_t n,
                                                                  uint32_t* __restrict__ sums)
{
  __shared__ uint32_t sm[SCAN_THREADS / 32];
  int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK;
  uint32_t acc = 0;
#pragma unroll
  for (int r = 0; r < SCAN_ROUNDS; ++r) {
    uint4 v = load4(d, base + (int64_t)r * SCAN_THREADS * 4 + threadIdx.x * 4, n);
    acc += v.x + v.y + v.z + v.w;
  }
  uint32_t t = block_reduce(acc, sm);
  if (threadIdx.x == 0)
    sums[blockIdx.x] = t;
}

// exclusive scan of one chunk with a base offset; optionally emits base+chunk total
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(uint32_t* __restrict__ d, int64_t n,
                                                                const uint32_t* __restrict__ bases,
                                                                uint32_t* __restrict__ total_out)
{
  __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
  __shared__ uint32_t carry;
  int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0)
    carry = bases ? bases[blockIdx.x] : 0u;
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < SCAN_ROUNDS; ++r) {
    int64_t i = base + (int64_t)r * SCAN_THREADS * 4 + threadIdx.x * 4;
    uint4 v = load4(d, i, n);
    uint32_t tsum = v.x + v.y + v.z + v.w;
    uint32_t inc = tsum;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint32_t wbase = 0, all = 0;
    for (int k = 0; k < SCAN_THREADS / 32; ++k) {
      uint32_t t = warp_tot[k];
      if (k < w) wbase += t;
      all += t;
    }
    uint32_t ex = carry + wbase + inc - tsum;
    uint4 o4;
    o4.x = ex;
    o4.y = ex + v.x;
    o4.z = o4.y + v.y;
    o4.w = o4.z + v.z;
    if (i + 3 < n) {
      *reinterpret_cast<uint4*>(d + i) = o4;
    } else {
      if (i < n) d[i] = o4.x;
      if (i + 1 < n) d[i + 1] = o4.y;
      if (i + 2 < n) d[i + 2] = o4.z;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += all;
    __syncthreads();
  }
  if (total_out && threadIdx.x == 0 && blockIdx.x == gridDim.x - 1)
    *total_out = carry;
}

int scan_rec(bseg_ctx* c, uint32_t* d, int64_t n, uint32_t* d_total, uint32_t* scratch, size_t scratch_elems)
{
  if (n <= 0) {
    if (d_total) CU_CHECK(c, cudaMemsetAsync(d_total, 0, sizeof(uint32_t), c->stream));
    return 0;
  }
  int64_t nb = ceil_div64(n, SCAN_CHUNK);
  if (nb == 1) {
    scan_down_kernel<<<1, SCAN_THREADS, 0, c->stream>>>(d, n, nullptr, d_total);
    KLAUNCH_CHECK(c);
    return 0;
  }
  if ((size_t)nb > scratch_elems)
    return bseg_fail(c, BSEG_E_CAPACITY, "scan scratch too small");
  scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, c->stream>>>(d, n, scratch);
  KLAUNCH_CHECK(c);
  // pad so the recursive level's scratch starts 16-byte aligned
  int64_t used = (nb + 3) & ~int64_t(3);
  RC_CHECK(scan_rec(c, scratch, nb, nullptr, scratch + used, scratch_elems - (size_t)used));
  scan_down_kernel<<<(unsigned)nb, SCAN_THREADS, 0, c->stream>>>(d, n, scratch, d_total);
  KLAUNCH_CHECK(c);
  return 0;
}

}  // namespace

int bseg_exclusive_scan_u32(bseg_ctx* c, uint32_t* d_data, int64_t n, uint32_t* d_total)
{
  int64_t nb = ceil_div64(n > 0 ? n : 1, SCAN_CHUNK);
  size_t need = (size_t)(nb + nb / SCAN_CHUNK + 1024) * 2;
  RC_CHECK(dev_ensure(c, c->scan_tmp, need * sizeof(uint32_t)));
  return scan_rec(c, d_data, n, d_total, dptr<uint32_t>(c->scan_tmp), need);
}
