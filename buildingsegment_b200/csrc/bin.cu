// bin.cu -- stage a3 (bbox + shift, TMC3.cpp:58-72) and north_star stage (1): Morton / voxel binning.
//
// The reference's spatial index is Open3D's KD-tree, built twice (my_function.h:63,71); here the
// cloud is binned into cubic cells of edge `cell` >= radius by a radix sort on the Morton code of
// the cell coordinates.  After the sort
//   pts[s]        = (x, y, z, original index) of the s-th point in Morton order      16 B
//   inv[orig]     = s
//   cell_key[c]   = Morton code of occupied cell c (ascending), cell_start[c..c+1] its point range
//   hash          = open-addressing table cell_key -> c, for the 27-neighbourhood lookups
// Every level-L ancestor cell (edge cell << L) is a contiguous range of the sorted keys, which is
// what the kNN fallback uses for ring expansion.
//
// All kernels here are HBM-bound streaming passes; algorithmic bytes are given per kernel.
#include "common.cuh"

namespace {

constexpr int TPB = 256;

// ---- bbox: R 12 B/pt ------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) bbox_kernel(const int32_t* __restrict__ xyz, int64_t n, int32_t* __restrict__ mm)
{
  int32_t mn[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, mx[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
  // flat, fully coalesced walk over the AoS: element e belongs to axis e % 3
  int64_t total = n * 3;
  int64_t stride = (int64_t)gridDim.x * TPB;  // multiple of 3? not required: axis recomputed per element
  for (int64_t e = (int64_t)blockIdx.x * TPB + threadIdx.x; e < total; e += stride) {
    int32_t v = xyz[e];
    int a = (int)(e % 3);
    if (a == 0) { mn[0] = min(mn[0], v); mx[0] = max(mx[0], v); }
    else if (a == 1) { mn[1] = min(mn[1], v); mx[1] = max(mx[1], v); }
    else { mn[2] = min(mn[2], v); mx[2] = max(mx[2], v); }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mn[k] = __reduce_min_sync(FULL_MASK, mn[k]);
    mx[k] = __reduce_max_sync(FULL_MASK, mx[k]);
  }
  __shared__ int32_t s[TPB / 32][6];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < 3; ++k) { s[w][k] = mn[k]; s[w][3 + k] = mx[k]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    int32_t v = s[0][threadIdx.x];
    for (int i = 1; i < TPB / 32; ++i)
      v = threadIdx.x < 3 ? min(v, s[i][threadIdx.x]) : max(v, s[i][threadIdx.x]);
    if (threadIdx.x < 3) atomicMin(&mm[threadIdx.x], v);
    else atomicMax(&mm[threadIdx.x], v);
  }
}

// ---- shift: R 12 + W 12 B/pt (TMC3.cpp:70-72: p -= min, wrap-around like the int32 original) -------
__global__ void __launch_bounds__(TPB) shift_kernel(int32_t* __restrict__ xyz, int64_t n, const int32_t* __restrict__ mm)
{
  uint32_t m0 = (uint32_t)mm[0], m1 = (uint32_t)mm[1], m2 = (uint32_t)mm[2];
  int64_t total = n * 3;
  int64_t stride = (int64_t)gridDim.x * TPB;
  for (int64_t e = (int64_t)blockIdx.x * TPB + threadIdx.x; e < total; e += stride) {
    int a = (int)(e % 3);
    uint32_t m = a == 0 ? m0 : (a == 1 ? m1 : m2);
    xyz[e] = (int32_t)((uint32_t)xyz[e] - m);
  }
}

// ---- cell keys: R 12, W 8 + 4 B/pt --------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) key_kernel(const int32_t* __restrict__ xyz, int64_t n, int32_t cell,
                                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
  int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n)
    return;
  uint32_t x = (uint32_t)xyz[3 * i] / (uint32_t)cell;
  uint32_t y = (uint32_t)xyz[3 * i + 1] / (uint32_t)cell;
  uint32_t z = (uint32_t)xyz[3 * i + 2] / (uint32_t)cell;
  keys[i] = morton3(x, y, z);
  vals[i] = (uint32_t)i;
}

// ---- occupancy per level: hist[b] = #adjacent sorted pairs whose cells first differ at level b -----
// n_cells(L) = 1 + sum_{b >= L} hist[b].  R 8 B/pt.
__global__ void __launch_bounds__(TPB) level_hist_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                         uint32_t* __restrict__ hist)
{
  __shared__ uint32_t h[32];
  if (threadIdx.x < 32) h[threadIdx.x] = 0;
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i > 0 && i < n) {
    uint64_t d = keys[i] ^ keys[i - 1];
    if (d) {
      int b = (63 - __clzll((long long)d)) / 3;
      atomicAdd(&h[b], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32 && h[threadIdx.x])
    atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// ---- gather sorted points + inverse permutation + cell heads: R 8+4+12 (gather), W 16+4+4+8 B/pt --------
__global__ void __launch_bounds__(TPB) gather_kernel(const int32_t* __restrict__ xyz, uint64_t* __restrict__ keys,
                                                     const uint32_t* __restrict__ vals, int64_t n, int lvl3,
                                                     int4* __restrict__ pts, uint32_t* __restrict__ inv,
                                                     uint32_t* __restrict__ head)
{
  int64_t s = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (s >= n)
    return;
  uint32_t o = vals[s];
  uint64_t k = keys[s] >> lvl3;
  uint64_t kp = s > 0 ? (keys[s - 1] >> lvl3) : ~0ull;
  pts[s] = make_int4(xyz[3 * (int64_t)o], xyz[3 * (int64_t)o + 1], xyz[3 * (int64_t)o + 2], (int)o);
  inv[o] = (uint32_t)s;
  head[s] = (k != kp) ? 1u : 0u;
}

// heads of the cells at another level of the same sorted keys
__global__ void __launch_bounds__(TPB) head_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                   uint32_t* __restrict__ head)
{
  int64_t s = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (s >= n)
    return;
  head[s] = (s == 0 || (keys[s] >> shift) != (keys[s - 1] >> shift)) ? 1u : 0u;
}

// keys >>= lvl3 in a second pass (the gather reads its left neighbour's unshifted key)
__global__ void __launch_bounds__(TPB) shift_keys_kernel(uint64_t* __restrict__ keys, int64_t n, int lvl3)
{
  int64_t s = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (s < n)
    keys[s] >>= lvl3;
}

// ---- cell table from scanned heads: R 4+4+8, W 12 B/cell ---------------------------------------------
__global__ void __launch_bounds__(TPB) cells_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ excl,
                                                    int64_t n, uint64_t* __restrict__ cell_key,
                                                    uint32_t* __restrict__ cell_start, uint32_t n_cells, int shift = 0)
{
  int64_t s = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (s >= n)
    return;
  uint32_t e = excl[s];
  bool is_head = (s + 1 < n) ? (excl[s + 1] != e) : (e + 1 == n_cells);
  // head[s] == 1  <=>  excl[s+1] == excl[s] + 1; the last element is a head iff e == n_cells - 1
  if (is_head) {
    cell_key[e] = keys[s] >> shift;
    cell_start[e] = (uint32_t)s;
  }
  if (s == n - 1)
    cell_start[n_cells] = (uint32_t)n;
}

__global__ void __launch_bounds__(TPB) hash_build_kernel(const uint64_t* __restrict__ cell_key, uint32_t n_cells,
                                                         unsigned long long* __restrict__ hk, uint32_t* __restrict__ hv,
                                                         uint64_t mask)
{
  uint32_t c = blockIdx.x * TPB + threadIdx.x;
  if (c >= n_cells)
    return;
  uint64_t key = cell_key[c];
  uint64_t h = hash64(key) & mask;
  for (;;) {
    unsigned long long old = atomicCAS(&hk[h], (unsigned long long)HASH_EMPTY, (unsigned long long)key);
    if (old == HASH_EMPTY) {
      hv[h] = c;
      return;
    }
    h = (h + 1) & mask;
  }
}

int grid_for(bseg_ctx* c, int64_t work_items, int per_block)
{
  int64_t b = ceil_div64(work_items, per_block);
  int64_t cap = (int64_t)c->num_sms * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

int read_back(bseg_ctx* c, void* host_dst, const void* dev_src, size_t bytes)
{
  if (c->pinned_cap < bytes) {
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr;
    c->pinned_cap = 0;
    size_t want = bytes < 4096 ? 4096 : bytes;
    CU_CHECK(c, cudaMallocHost(&c->pinned, want));
    c->pinned_cap = want;
  }
  CU_CHECK(c, cudaMemcpyAsync(c->pinned, dev_src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  memcpy(host_dst, c->pinned, bytes);
  return 0;
}

int stage_bbox_shift(bseg_ctx* c)
{
  RC_CHECK(dev_ensure(c, c->minmax, 8 * sizeof(int32_t)));
  if (c->n == 0) {
    for (int k = 0; k < 3; ++k) { c->mn[k] = 0; c->mx[k] = 0; }
    return 0;
  }
  STAGE_BEGIN(c, EV_BBOX);
  int32_t init[8] = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN, INT32_MIN, 0, 0};
  CU_CHECK(c, cudaMemcpyAsync(c->minmax.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  int g = grid_for(c, c->n * 3, TPB * 8);
  bbox_kernel<<<g, TPB, 0, c->stream>>>(dptr<int32_t>(c->xyz_raw), c->n, dptr<int32_t>(c->minmax));
  KLAUNCH_CHECK(c);
  int32_t mm[8];
  if (c->have_origin) {
    // multi-GPU slabs share the tile's origin: shift by it instead of this slab's own minimum
    RC_CHECK(read_back(c, mm, c->minmax.p, sizeof(mm)));
    for (int k = 0; k < 3; ++k)
      if (c->origin[k] > mm[k])
        return bseg_fail(c, BSEG_E_ARG, "bseg_set_origin: origin[%d] = %d lies above the cloud's minimum %d", k,
                         c->origin[k], mm[k]);
    CU_CHECK(c, cudaMemcpyAsync(c->minmax.p, c->origin, 3 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  }
  shift_kernel<<<g, TPB, 0, c->stream>>>(dptr<int32_t>(c->xyz_raw), c->n, dptr<int32_t>(c->minmax));
  KLAUNCH_CHECK(c);
  STAGE_END(c, EV_BBOX);
  RC_CHECK(read_back(c, mm, c->minmax.p, sizeof(mm)));
  for (int k = 0; k < 3; ++k) { c->mn[k] = mm[k]; c->mx[k] = mm[3 + k]; }  // mn = the origin that was subtracted
  return 0;
}

static int bits_for(uint32_t v)
{
  int b = 0;
  while (v) { ++b; v >>= 1; }
  return b;
}

int stage_bin(bseg_ctx* c, const bseg_params* p)
{
  const int64_t n = c->n;
  c->K = p->K;
  int32_t c0 = (int32_t)p->radius;
  if ((double)c0 < p->radius) ++c0;  // ceil
  if (c0 < 1) c0 = 1;
  if (p->cell > 0) c0 = p->cell;
  c->cell = c0;
  c->n_cells = 0;
  if (n == 0)
    return 0;
  for (int i = 0; i < 2; ++i) {
    RC_CHECK(dev_ensure(c, c->keys[i], (size_t)n * 8));
    RC_CHECK(dev_ensure(c, c->vals[i], (size_t)n * 4));
  }
  RC_CHECK(dev_ensure(c, c->pts, (size_t)n * 16));
  RC_CHECK(dev_ensure(c, c->inv, (size_t)n * 4));
  RC_CHECK(dev_ensure(c, c->flags, (size_t)(n + 4) * 4));
  RC_CHECK(dev_ensure(c, c->counters, 192 * sizeof(uint64_t)));

  uint32_t ext[3];
  int maxbits = 0;
  for (int k = 0; k < 3; ++k) {
    ext[k] = (uint32_t)(c->mx[k] - c->mn[k]);
    int b = bits_for(ext[k] / (uint32_t)c0);
    if (b > maxbits) maxbits = b;
  }
  if (maxbits > 21)
    return bseg_fail(c, BSEG_E_ARG, "more than 2^21 kNN cells along one axis");
  c->key_bits = 3 * (maxbits > 0 ? maxbits : 1);

  const unsigned nb = (unsigned)ceil_div64(n, TPB);
  STAGE_BEGIN(c, EV_SORT);
  key_kernel<<<nb, TPB, 0, c->stream>>>(dptr<int32_t>(c->xyz_raw), n, c0, dptr<uint64_t>(c->keys[0]),
                                        dptr<uint32_t>(c->vals[0]));
  KLAUNCH_CHECK(c);
  int sel = 0;
  RC_CHECK(bseg_sort_pairs_u64(c, dptr<uint64_t>(c->keys[0]), dptr<uint64_t>(c->keys[1]), dptr<uint32_t>(c->vals[0]),
                               dptr<uint32_t>(c->vals[1]), n, c->key_bits, &sel));
  c->sort_sel = sel;
  STAGE_END(c, EV_SORT);

  STAGE_BEGIN(c, EV_CELLS);
  uint64_t* keys = dptr<uint64_t>(c->keys[sel]);
  uint32_t* vals = dptr<uint32_t>(c->vals[sel]);
  uint32_t* hist = reinterpret_cast<uint32_t*>(dptr<uint64_t>(c->counters));
  CU_CHECK(c, cudaMemsetAsync(hist, 0, 32 * sizeof(uint32_t), c->stream));
  level_hist_kernel<<<nb, TPB, 0, c->stream>>>(keys, n, hist);
  KLAUNCH_CHECK(c);
  uint32_t h[32];
  RC_CHECK(read_back(c, h, hist, sizeof(h)));
  // choose the level: the coarsest-needed cell so that an occupied cell holds ~>= 10 points on
  // average (the K-th neighbour then usually lies inside the 27-cell block), cell <= BSEG_MAX_CELL
  int lvl = 0;
  int64_t cells_at[23];
  for (int L = 22; L >= 0; --L)
    cells_at[L] = (L == 22 ? 1 : cells_at[L + 1]) + (L < 32 ? h[L] : 0);
  // cells_at[L] = 1 + sum_{b >= L} h[b]
  if (p->cell == 0) {
    const double target = (double)p->K * 0.7;
    while (lvl < 20 && (int64_t)c0 << (lvl + 1) <= BSEG_MAX_CELL && (double)n / (double)cells_at[lvl] < target)
      ++lvl;
  }
  c->cell = c0 << lvl;
  c->n_cells = cells_at[lvl];
  c->key_bits -= 3 * lvl;
  if (c->key_bits < 3) c->key_bits = 3;
  const int lvl3 = 3 * lvl;
  uint32_t* head = dptr<uint32_t>(c->flags);
  gather_kernel<<<nb, TPB, 0, c->stream>>>(dptr<int32_t>(c->xyz_raw), keys, vals, n, lvl3, dptr<int4>(c->pts),
                                           dptr<uint32_t>(c->inv), head);
  KLAUNCH_CHECK(c);
  // the finer table (edge cell / 2) for the dense groups of the kNN: same sorted cloud, one level down
  c->have_mid = false;
  if (lvl >= 1) {
    const int shift2 = 3 * (lvl - 1);
    c->n_cells2 = cells_at[lvl - 1];
    uint32_t* head2 = dptr<uint32_t>(c->vals[1 - sel]);  // the sort's other value buffer is free now
    head_kernel<<<nb, TPB, 0, c->stream>>>(keys, n, shift2, head2);
    KLAUNCH_CHECK(c);
    RC_CHECK(bseg_exclusive_scan_u32(c, head2, n, nullptr));
    RC_CHECK(dev_ensure(c, c->cell_key2, (size_t)c->n_cells2 * 8));
    RC_CHECK(dev_ensure(c, c->cell_start2, (size_t)(c->n_cells2 + 1) * 4));
    cells_kernel<<<nb, TPB, 0, c->stream>>>(keys, head2, n, dptr<uint64_t>(c->cell_key2), dptr<uint32_t>(c->cell_start2),
                                            (uint32_t)c->n_cells2, shift2);
    KLAUNCH_CHECK(c);
    uint64_t hsize2 = 1024;
    while (hsize2 < (uint64_t)c->n_cells2 * 2) hsize2 <<= 1;
    c->hash_mask2 = hsize2 - 1;
    RC_CHECK(dev_ensure(c, c->hash_keys2, hsize2 * 8));
    RC_CHECK(dev_ensure(c, c->hash_vals2, hsize2 * 4));
    CU_CHECK(c, cudaMemsetAsync(c->hash_keys2.p, 0xff, hsize2 * 8, c->stream));
    hash_build_kernel<<<(unsigned)ceil_div64(c->n_cells2, TPB), TPB, 0, c->stream>>>(
        dptr<uint64_t>(c->cell_key2), (uint32_t)c->n_cells2, reinterpret_cast<unsigned long long*>(c->hash_keys2.p),
        dptr<uint32_t>(c->hash_vals2), c->hash_mask2);
    KLAUNCH_CHECK(c);
    c->have_mid = true;
  }
  if (lvl3) {
    shift_keys_kernel<<<nb, TPB, 0, c->stream>>>(keys, n, lvl3);
    KLAUNCH_CHECK(c);
  }
  RC_CHECK(bseg_exclusive_scan_u32(c, head, n, nullptr));
  RC_CHECK(dev_ensure(c, c->cell_key, (size_t)c->n_cells * 8));
  RC_CHECK(dev_ensure(c, c->cell_start, (size_t)(c->n_cells + 1) * 4));
  cells_kernel<<<nb, TPB, 0, c->stream>>>(keys, head, n, dptr<uint64_t>(c->cell_key), dptr<uint32_t>(c->cell_start),
                                          (uint32_t)c->n_cells);
  KLAUNCH_CHECK(c);
  uint64_t hsize = 1024;
  while (hsize < (uint64_t)c->n_cells * 2) hsize <<= 1;
  c->hash_mask = hsize - 1;
  RC_CHECK(dev_ensure(c, c->hash_keys, hsize * 8));
  RC_CHECK(dev_ensure(c, c->hash_vals, hsize * 4));
  CU_CHECK(c, cudaMemsetAsync(c->hash_keys.p, 0xff, hsize * 8, c->stream));
  hash_build_kernel<<<(unsigned)ceil_div64(c->n_cells, TPB), TPB, 0, c->stream>>>(
      dptr<uint64_t>(c->cell_key), (uint32_t)c->n_cells, reinterpret_cast<unsigned long long*>(c->hash_keys.p),
      dptr<uint32_t>(c->hash_vals), c->hash_mask);
  KLAUNCH_CHECK(c);
  STAGE_END(c, EV_CELLS);
  return 0;
}
