// knn.cu -- north_star stages (2) + (3): exact k-nearest-neighbour / fixed-radius gather over the
// voxel grid with the PCA normal fused into the epilogue.
//
// Replaces get_Normal_and_K_neighbor<K> (my_function.h:48-85): Open3D EstimateNormals with
// KDTreeSearchParamHybrid(radius, max_nn) + OrientNormalsToAlignWithDirection((0,0,1)) (:63-64)
// and KDTreeFlann::SearchKNN(p, K) per point (:75-78).  Both neighbour sets are prefixes of ONE
// total order, (d^2 ascending, original index ascending), d^2 exact in integers:
//   row(i)    = first K entries                                   (self included)
//   hybrid(i) = first min(max_nn, #{d^2 < radius^2}) entries      -> covariance -> normal
//
// One warp owns one occupied cell.  It looks its 27 neighbour cells up in the hash table, stages
// their points in shared memory (x,y,z,orig + sorted position), then serves the cell's queries
// one after another: every lane evaluates M/32 candidates, the K-th / max_nn-th smallest
// (d^2, idx) is found by a warp-wide bisection on d^2 (count with one REDUX per probe) and, only
// when ties straddle the cut, a second bisection on the index.  The selected candidates'
// integer moments are reduced over the warp; once 32 queries are done each lane runs the
// closed-form fp64 eigen-solver (bseg_arith.h) for one of them.
//
// A row is exact when its K-th distance is below cell+1 (nothing outside the 27 cells can be
// closer); other queries go to the ring-expansion fallback (level-L ancestor cells are contiguous
// key ranges of the Morton-sorted cloud).  The hybrid set is always exact because cell >= radius.
#include <cstdlib>

#include "common.cuh"
#include "bseg_arith.h"

namespace {

constexpr int KW = 4;                  // warps per block
constexpr int KTHREADS = KW * 32;
constexpr int CAP = 512;               // staged candidates per warp

struct KnnArgs {
  const int4* pts;
  const uint64_t* keys;       // sorted cell keys, one per point
  const uint32_t* cell_start;
  const uint64_t* cell_key;
  uint32_t n_cells;
  const uint64_t* hk;
  const uint32_t* hv;
  uint64_t hmask;
  int32_t cell;
  uint32_t gmax[3];           // largest cell coordinate per axis
  int K, max_nn;
  uint32_t r2i;               // d2 < radius^2  <=>  d2 < r2i
  uint32_t guar2;             // (cell+1)^2: rows whose K-th d2 is below are exact
  int32_t* nbr;
  double* nrm;
  double* curv;
  uint32_t* big_cells;        // work items (cell, batch of 32 queries) of cells whose block exceeds CAP
  uint32_t* n_big;
  uint32_t* unres;            // queries needing ring expansion
  uint32_t* n_unres;
  int64_t n;
  const uint32_t* inv;        // original index -> sorted position
  // cells the group kernel left to the per-cell kernels: cell | HYB_ONLY (rows done, hybrid set still to do)
  uint32_t* cell_list;
  uint32_t* n_cell_list;
  uint32_t* next_chunk;       // work counter of the group kernel
  // 32-query batches of groups whose block does not fit the staging area (streamed in chunks, one warp per batch)
  uint32_t* dense_items;      // (cell | level << 31, batch)
  uint32_t* n_dense;
  uint32_t* next_dense;
};
constexpr uint32_t HYB_ONLY = 0x80000000u;

// selection predicate: d2 < t, plus (eq && d2 == t && idx <= u)
struct Sel {
  uint32_t t, u;
  int eq;
  __device__ __forceinline__ bool take(uint32_t d2, uint32_t idx) const
  {
    return d2 < t || (eq && d2 == t && idx <= u);
  }
};

__device__ __forceinline__ uint32_t dist2(const int4& a, const int4& q)
{
  int dx = a.x - q.x, dy = a.y - q.y, dz = a.z - q.z;
  return (uint32_t)(dx * dx) + (uint32_t)(dy * dy) + (uint32_t)(dz * dz);
}

// ---- candidate views --------------------------------------------------------------------------------
// staged in shared memory, d2 cached in R registers per lane
template <int R>
struct CachedView {
  uint32_t d2[R];
  const int4* sp;
  const uint32_t* spos;
  int M, lane;
  __device__ __forceinline__ void load(const int4& q)
  {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int j = r * 32 + lane;
      d2[r] = j < M ? dist2(sp[j], q) : 0xffffffffu;
    }
  }
  template <class F>
  __device__ __forceinline__ void for_each(F&& f) const
  {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int j = r * 32 + lane;
      if (j < M)
        f(d2[r], sp[j], spos[j]);
    }
  }
  template <class F>
  __device__ __forceinline__ void for_each_d2(F&& f) const
  {
#pragma unroll
    for (int r = 0; r < R; ++r)
      f(d2[r]);
  }
};

// streamed from global memory through the 27 ranges, d2 recomputed on every pass
struct GlobalView {
  const int4* pts;
  const uint32_t* rs;  // range starts (shared)
  const uint32_t* rl;  // range lengths (shared)
  int4 q;
  int lane;
  __device__ __forceinline__ void load(const int4& qq) { q = qq; }
  template <class F>
  __device__ __forceinline__ void for_each(F&& f) const
  {
    for (int c = 0; c < 27; ++c) {
      uint32_t s = rs[c], l = rl[c];
      for (uint32_t t = lane; t < l; t += 32) {
        int4 p = __ldg(pts + s + t);
        f(dist2(p, q), p, s + t);
      }
    }
  }
  template <class F>
  __device__ __forceinline__ void for_each_d2(F&& f) const
  {
    for_each([&](uint32_t d2, const int4&, uint32_t) { f(d2); });
  }
};

template <class V>
__device__ __forceinline__ uint32_t count_le(const V& v, uint32_t t)
{
  uint32_t c = 0;
  v.for_each_d2([&](uint32_t d2) { c += (d2 <= t) ? 1u : 0u; });
  return __reduce_add_sync(FULL_MASK, c);
}

// k smallest by (d2, idx) among candidates with d2 <= hi; precondition count_le(hi) >= k, hi < 2^32-1.
// Finds t* = min{t : #{d2 <= t} >= k} by a bracketed search that alternates an interpolation probe
// (#{d2 <= t} grows about linearly in t on a surface: area ~ r^2 = t) with a bisection probe; a probe
// that lands between the k-th and (k+1)-th distance ends the search at once.
template <class V>
__device__ Sel select_k(const V& v, uint32_t k, uint32_t hi, uint32_t idx_hi)
{
  uint32_t lo = 0;
  uint32_t flo = 0;            // #{d2 <= lo - 1}  (< k)
  uint32_t fhi = 0xffffffffu;  // #{d2 <= hi}, unknown until probed
  bool interp = false;
  while (lo < hi) {
    uint32_t mid;
    if (interp && fhi != 0xffffffffu && fhi > flo) {
      const uint64_t span = (uint64_t)(hi - lo);
      mid = lo + (uint32_t)((span * (uint64_t)(k - flo)) / (uint64_t)(fhi - flo));
      if (mid >= hi) mid = hi - 1;
    } else {
      mid = lo + ((hi - lo) >> 1);
    }
    interp = !interp;
    const uint32_t c = count_le(v, mid);
    if (c == k)
      return Sel{mid + 1, 0, 0};
    if (c > k) {
      hi = mid;
      fhi = c;
    } else {
      lo = mid + 1;
      flo = c;
    }
  }
  const uint32_t t = lo;
  uint32_t less = 0, ties = 0;
  v.for_each_d2([&](uint32_t d2) {
    less += (d2 < t) ? 1u : 0u;
    ties += (d2 == t) ? 1u : 0u;
  });
  less = __reduce_add_sync(FULL_MASK, less);
  ties = __reduce_add_sync(FULL_MASK, ties);
  const uint32_t need = k - less;
  if (ties == need)
    return Sel{t + 1, 0, 0};
  // ties straddle the cut: the `need` smallest original indices among d2 == t
  uint32_t ulo = 0, uhi = idx_hi;
  while (ulo < uhi) {
    uint32_t mid = ulo + ((uhi - ulo) >> 1);
    uint32_t c = 0;
    v.for_each([&](uint32_t d2, const int4& p, uint32_t) { c += (d2 == t && (uint32_t)p.w <= mid) ? 1u : 0u; });
    c = __reduce_add_sync(FULL_MASK, c);
    if (c >= need) uhi = mid;
    else ulo = mid + 1;
  }
  return Sel{t, ulo, 1};
}

struct WarpScratch {
  uint32_t rs[27], rl[27];
  uint32_t sel_d2[32], sel_idx[32], sel_pos[32];
};

// K-row of one query from a candidate view holding `M` candidates.  `complete` says the view is
// known to contain every point with d2 < A.guar2 (true for the 27-cell block).  Returns false when
// the row could not be proven exact (left to the ring-expansion fallback).
template <class V>
__device__ __forceinline__ bool row_from_view(const KnnArgs& A, const V& v, WarpScratch* ws, uint32_t qpos, uint32_t M,
                                              int lane)
{
  const uint32_t K = (uint32_t)A.K;
  uint32_t nsel = 0;
  bool resolved = false;
  if (M >= K) {
    uint32_t mx = 0;
    v.for_each_d2([&](uint32_t d2) { mx = (d2 != 0xffffffffu && d2 > mx) ? d2 : mx; });
    mx = __reduce_max_sync(FULL_MASK, mx);
    const Sel s = select_k(v, K, mx, (uint32_t)(A.n - 1));
    uint32_t kth = 0, mine = 0;
    v.for_each([&](uint32_t d2, const int4& p, uint32_t) {
      if (s.take(d2, (uint32_t)p.w)) {
        ++mine;
        if (d2 > kth) kth = d2;
      }
    });
    kth = __reduce_max_sync(FULL_MASK, kth);
    resolved = kth < A.guar2;
    // compact the K selected into scratch (per-lane serial slots, warp-wide exclusive offsets)
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    uint32_t w = incl - mine;
    v.for_each([&](uint32_t d2, const int4& p, uint32_t pos) {
      if (s.take(d2, (uint32_t)p.w)) {
        if (w < 32) {
          ws->sel_d2[w] = d2;
          ws->sel_idx[w] = (uint32_t)p.w;
          ws->sel_pos[w] = pos;
        }
        ++w;
      }
    });
    nsel = K;
  } else {
    // fewer candidates than K: take them all, leave the row to the fallback
    uint32_t mine = 0;
    v.for_each([&](uint32_t, const int4&, uint32_t) { ++mine; });
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    uint32_t w = incl - mine;
    v.for_each([&](uint32_t d2, const int4& p, uint32_t pos) {
      if (w < 32) {
        ws->sel_d2[w] = d2;
        ws->sel_idx[w] = (uint32_t)p.w;
        ws->sel_pos[w] = pos;
      }
      ++w;
    });
    nsel = M;
  }
  __syncwarp();
  int32_t* row = A.nbr + (int64_t)qpos * A.K;
  if ((uint32_t)lane < nsel) {
    uint32_t md = ws->sel_d2[lane], mi = ws->sel_idx[lane];
    uint32_t rank = 0;
    for (uint32_t m = 0; m < nsel; ++m) {
      uint32_t od = ws->sel_d2[m], oi = ws->sel_idx[m];
      rank += (od < md || (od == md && oi < mi)) ? 1u : 0u;
    }
    row[rank] = (int32_t)ws->sel_pos[lane];
  } else if (lane < A.K) {
    row[lane] = -1;
  }
  __syncwarp();
  return resolved;
}

__device__ __forceinline__ void push_unresolved(const KnnArgs& A, uint32_t qpos, int lane)
{
  if (lane == 0) {
    uint32_t slot = atomicAdd(A.n_unres, 1u);
    A.unres[slot] = qpos;
  }
}

// hybrid set of one query -> integer moments relative to `org`, on every lane
template <class V>
__device__ __forceinline__ void hybrid_from_view(const KnnArgs& A, const V& v, const int3& org, int& cnt_out,
                                                 int (&ms)[9])
{
  uint32_t cin = 0;
  v.for_each_d2([&](uint32_t d2) { cin += (d2 < A.r2i) ? 1u : 0u; });
  cin = __reduce_add_sync(FULL_MASK, cin);
  Sel h{A.r2i, 0, 0};
  uint32_t cnt = cin;
  if (cin > (uint32_t)A.max_nn) {
    h = select_k(v, (uint32_t)A.max_nn, A.r2i - 1, (uint32_t)(A.n - 1));
    cnt = (uint32_t)A.max_nn;
  }
  int a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  v.for_each([&](uint32_t d2, const int4& p, uint32_t) {
    if (h.take(d2, (uint32_t)p.w)) {
      int x = p.x - org.x, y = p.y - org.y, z = p.z - org.z;
      a[0] += x; a[1] += y; a[2] += z;
      a[3] += x * x; a[4] += x * y; a[5] += x * z;
      a[6] += y * y; a[7] += y * z; a[8] += z * z;
    }
  });
#pragma unroll
  for (int m = 0; m < 9; ++m)
    ms[m] = __reduce_add_sync(FULL_MASK, a[m]);
  cnt_out = (int)cnt;
}

// One query against the staged 27-cell block: K-row + hybrid moments.
template <class V>
__device__ __forceinline__ void serve_query(const KnnArgs& A, V& v, WarpScratch* ws, const int4& q, uint32_t qpos,
                                            uint32_t M, const int3& org, int lane, int& cnt_out, int (&ms)[9], bool rows)
{
  v.load(q);
  if (rows && !row_from_view(A, v, ws, qpos, M, lane))
    push_unresolved(A, qpos, lane);
  hybrid_from_view(A, v, org, cnt_out, ms);
}

// per-lane epilogue: relative integer moments -> absolute exact sums -> fp64 normal (+ curvature)
__device__ __forceinline__ void finish_normal(const KnnArgs& A, uint32_t qpos, int cnt, const int (&ms)[9],
                                              const int3& org)
{
  const int64_t n = cnt, ox = org.x, oy = org.y, oz = org.z;
  const int64_t sx = ms[0], sy = ms[1], sz = ms[2];
  double sums[9];
  sums[0] = (double)(sx + n * ox);
  sums[1] = (double)(sy + n * oy);
  sums[2] = (double)(sz + n * oz);
  sums[3] = (double)((int64_t)ms[3] + 2 * ox * sx + n * ox * ox);
  sums[4] = (double)((int64_t)ms[4] + ox * sy + oy * sx + n * ox * oy);
  sums[5] = (double)((int64_t)ms[5] + ox * sz + oz * sx + n * ox * oz);
  sums[6] = (double)((int64_t)ms[6] + 2 * oy * sy + n * oy * oy);
  sums[7] = (double)((int64_t)ms[7] + oy * sz + oz * sy + n * oy * oz);
  sums[8] = (double)((int64_t)ms[8] + 2 * oz * sz + n * oz * oz);
  double nr[3], cv;
  bseg_normal_from_sums(sums, cnt, nr, &cv);
  double* o = A.nrm + (int64_t)qpos * 3;
  o[0] = nr[0];
  o[1] = nr[1];
  o[2] = nr[2];
  A.curv[qpos] = cv;
}

template <class V>
__device__ __forceinline__ void serve_cell(const KnnArgs& A, V& v, WarpScratch* ws, uint32_t qstart, uint32_t qlen,
                                           uint32_t M, const int3& org, int lane, bool rows)
{
  for (uint32_t b0 = 0; b0 < qlen; b0 += 32) {
    const uint32_t nb = min(32u, qlen - b0);
    int my_cnt = 0;
    int my_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t qi = 0; qi < nb; ++qi) {
      const uint32_t qpos = qstart + b0 + qi;
      const int4 q = __ldg(A.pts + qpos);
      int cnt;
      int ms[9];
      serve_query(A, v, ws, q, qpos, M, org, lane, cnt, ms, rows);
      if ((uint32_t)lane == qi) {
        my_cnt = cnt;
#pragma unroll
        for (int m = 0; m < 9; ++m) my_ms[m] = ms[m];
      }
    }
    if ((uint32_t)lane < nb)
      finish_normal(A, qstart + b0 + lane, my_cnt, my_ms, org);
  }
}

// locate the 27 neighbour cells of `cell`; fills ws->rs/rl, returns the total candidate count
__device__ __forceinline__ uint32_t find_ranges(const KnnArgs& A, uint32_t cell, WarpScratch* ws, int lane, int3& org)
{
  const uint64_t key = __ldg(A.cell_key + cell);
  const uint32_t cx = morton_compact21(key), cy = morton_compact21(key >> 1), cz = morton_compact21(key >> 2);
  org = make_int3((int)cx * A.cell, (int)cy * A.cell, (int)cz * A.cell);
  uint32_t start = 0, len = 0;
  if (lane < 27) {
    int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
    int64_t nx = (int64_t)cx + dx, ny = (int64_t)cy + dy, nz = (int64_t)cz + dz;
    if (nx >= 0 && ny >= 0 && nz >= 0 && nx <= A.gmax[0] && ny <= A.gmax[1] && nz <= A.gmax[2]) {
      uint32_t c = (dx == 0 && dy == 0 && dz == 0)
                       ? cell
                       : hash_lookup(A.hk, A.hv, A.hmask, morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
      if (c != 0xffffffffu) {
        start = __ldg(A.cell_start + c);
        len = __ldg(A.cell_start + c + 1) - start;
      }
    }
    ws->rs[lane] = start;
    ws->rl[lane] = len;
  }
  uint32_t tot = __reduce_add_sync(FULL_MASK, len);
  __syncwarp();
  return tot;
}

// Per-cell kernel: the whole cloud when `use_list` is 0, else the cells the group kernel left over.
__global__ void __launch_bounds__(KTHREADS) knn_cells_kernel(KnnArgs A, int use_list)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int4* sp = reinterpret_cast<int4*>(smem) + (size_t)w * CAP;
  uint32_t* spos = reinterpret_cast<uint32_t*>(smem + (size_t)KW * CAP * 16) + (size_t)w * CAP;
  WarpScratch* ws = reinterpret_cast<WarpScratch*>(smem + (size_t)KW * CAP * 20) + w;

  const uint32_t nitems = use_list ? *A.n_cell_list : A.n_cells;
  for (uint32_t it = blockIdx.x * KW + w; it < nitems; it += gridDim.x * KW) {
    const uint32_t item = use_list ? A.cell_list[it] : it;
    const uint32_t cell = item & ~HYB_ONLY;
    const bool rows = (item & HYB_ONLY) == 0;
    int3 org;
    const uint32_t M = find_ranges(A, cell, ws, lane, org);
    const uint32_t qstart = __ldg(A.cell_start + cell);
    const uint32_t qlen = __ldg(A.cell_start + cell + 1) - qstart;
    if (M > CAP) {
      const uint32_t nbatch = (qlen + 31) / 32;
      uint32_t slot = 0;
      if (lane == 0) slot = atomicAdd(A.n_big, nbatch);
      slot = __shfl_sync(FULL_MASK, slot, 0);
      for (uint32_t b = lane; b < nbatch; b += 32) {
        A.big_cells[2 * (slot + b)] = item;
        A.big_cells[2 * (slot + b) + 1] = b;
      }
      continue;
    }
    // stage the candidates
    uint32_t off = 0;
    for (int c = 0; c < 27; ++c) {
      const uint32_t s = ws->rs[c], l = ws->rl[c];
      for (uint32_t t = lane; t < l; t += 32) {
        sp[off + t] = __ldg(A.pts + s + t);
        spos[off + t] = s + t;
      }
      off += l;
    }
    __syncwarp();
    if (M <= 128) {
      CachedView<4> v;
      v.sp = sp; v.spos = spos; v.M = (int)M; v.lane = lane;
      serve_cell(A, v, ws, qstart, qlen, M, org, lane, rows);
    } else if (M <= 256) {
      CachedView<8> v;
      v.sp = sp; v.spos = spos; v.M = (int)M; v.lane = lane;
      serve_cell(A, v, ws, qstart, qlen, M, org, lane, rows);
    } else {
      CachedView<16> v;
      v.sp = sp; v.spos = spos; v.M = (int)M; v.lane = lane;
      serve_cell(A, v, ws, qstart, qlen, M, org, lane, rows);
    }
    __syncwarp();
  }
}

// ---- dense neighbourhoods (> CAP candidates) ------------------------------------------------------
// Work item = (cell, batch of 32 queries).  Per query the warp streams the 27 ranges from global
// memory (L1/L2 resident: every query of the batch reads the same ranges) only to COUNT, finds a
// squared-distance threshold T that keeps between k and CAP candidates, compacts those into shared
// memory and runs the same cached selection as the regular path on them.
struct StreamCount {
  uint32_t a, b;
};

// number of candidates with d2 < ta and d2 < tb (one streaming pass)
__device__ __forceinline__ StreamCount stream_count(const GlobalView& gv, uint32_t ta, uint32_t tb)
{
  uint32_t ca = 0, cb = 0;
  gv.for_each_d2([&](uint32_t d2) {
    ca += (d2 < ta) ? 1u : 0u;
    cb += (d2 < tb) ? 1u : 0u;
  });
  StreamCount r;
  r.a = __reduce_add_sync(FULL_MASK, ca);
  r.b = __reduce_add_sync(FULL_MASK, cb);
  return r;
}

// Find T <= limit with  k <= #{d2 < T} <= CAP  (or T == limit when even that holds fewer than CAP).
// c_limit = #{d2 < limit} is known.  Returns 0 when no such T exists (more than CAP-k ties at one
// distance): the caller then falls back to the streaming selection.
__device__ uint32_t find_threshold(const GlobalView& gv, uint32_t limit, uint32_t c_limit, uint32_t k, uint32_t& c_out)
{
  if (c_limit <= (uint32_t)CAP) {
    c_out = c_limit;
    return limit;
  }
  uint32_t lo = 0, hi = limit;  // #{d2 < lo} < k,  #{d2 < hi} > CAP
  while (hi - lo > 1) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    const uint32_t c = stream_count(gv, mid, mid).a;
    if (c < k) lo = mid;
    else if (c > (uint32_t)CAP) hi = mid;
    else {
      c_out = c;
      return mid;
    }
  }
  return 0;
}

// compact the candidates with d2 < T into the warp's staging area; returns how many
__device__ __forceinline__ uint32_t stream_compact(const GlobalView& gv, uint32_t T, int4* sp, uint32_t* spos, int lane)
{
  uint32_t base = 0;
  for (int c = 0; c < 27; ++c) {
    const uint32_t s = gv.rs[c], l = gv.rl[c];
    for (uint32_t t0 = 0; t0 < l; t0 += 32) {
      const uint32_t t = t0 + lane;
      int4 p = make_int4(0, 0, 0, 0);
      bool in = false;
      if (t < l) {
        p = __ldg(gv.pts + s + t);
        in = dist2(p, gv.q) < T;
      }
      const uint32_t b = __ballot_sync(FULL_MASK, in);
      if (in) {
        const uint32_t w = base + __popc(b & lanemask_lt());
        sp[w] = p;
        spos[w] = s + t;
      }
      base += __popc(b);
    }
  }
  __syncwarp();
  return base;
}

__global__ void __launch_bounds__(KTHREADS) knn_big_cells_kernel(KnnArgs A)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int4* sp = reinterpret_cast<int4*>(smem) + (size_t)w * CAP;
  uint32_t* spos = reinterpret_cast<uint32_t*>(smem + (size_t)KW * CAP * 16) + (size_t)w * CAP;
  WarpScratch* ws = reinterpret_cast<WarpScratch*>(smem + (size_t)KW * CAP * 20) + w;
  const uint32_t nitems = *A.n_big;
  for (uint32_t it = blockIdx.x * KW + w; it < nitems; it += gridDim.x * KW) {
    const uint32_t item = A.big_cells[2 * it], batch = A.big_cells[2 * it + 1];
    const uint32_t cell = item & ~HYB_ONLY;
    const bool rows = (item & HYB_ONLY) == 0;
    int3 org;
    const uint32_t M = find_ranges(A, cell, ws, lane, org);
    const uint32_t cstart = __ldg(A.cell_start + cell);
    const uint32_t clen = __ldg(A.cell_start + cell + 1) - cstart;
    const uint32_t qstart = cstart + batch * 32;
    const uint32_t nb = min(32u, clen - batch * 32);
    GlobalView gv;
    gv.pts = A.pts; gv.rs = ws->rs; gv.rl = ws->rl; gv.lane = lane;
    int my_cnt = 0;
    int my_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t qi = 0; qi < nb; ++qi) {
      const uint32_t qpos = qstart + qi;
      const int4 q = __ldg(A.pts + qpos);
      gv.load(q);
      int cnt;
      int ms[9];
      const StreamCount sc = stream_count(gv, A.r2i, A.guar2);
      // ---- hybrid set: candidates with d2 < r2i, the max_nn nearest of them ----
      uint32_t c1 = 0;
      const uint32_t T1 = find_threshold(gv, A.r2i, sc.a, (uint32_t)A.max_nn, c1);
      bool row_done = !rows;
      if (T1) {
        const uint32_t m1 = stream_compact(gv, T1, sp, spos, lane);
        CachedView<16> v;
        v.sp = sp; v.spos = spos; v.M = (int)m1; v.lane = lane;
        v.load(q);
        hybrid_from_view(A, v, org, cnt, ms);
        if (rows && m1 >= (uint32_t)A.K) {  // the K nearest overall are among them (T1 <= r2i <= guar2)
          row_from_view(A, v, ws, qpos, m1, lane);
          row_done = true;
        }
      } else {
        hybrid_from_view(A, gv, org, cnt, ms);  // > CAP ties at one distance: exact streaming selection
      }
      // ---- K-row when fewer than K points lie inside the radius ----
      if (!row_done) {
        if (sc.b < (uint32_t)A.K) {
          if (lane < A.K) A.nbr[(int64_t)qpos * A.K + lane] = -1;
          push_unresolved(A, qpos, lane);  // the K-th neighbour may be outside the block
        } else {
          uint32_t c2 = 0;
          const uint32_t T2 = find_threshold(gv, A.guar2, sc.b, (uint32_t)A.K, c2);
          if (T2) {
            const uint32_t m2 = stream_compact(gv, T2, sp, spos, lane);
            CachedView<16> v;
            v.sp = sp; v.spos = spos; v.M = (int)m2; v.lane = lane;
            v.load(q);
            row_from_view(A, v, ws, qpos, m2, lane);
          } else if (!row_from_view(A, gv, ws, qpos, M, lane)) {
            push_unresolved(A, qpos, lane);
          }
        }
      }
      __syncwarp();
      if ((uint32_t)lane == qi) {
        my_cnt = cnt;
#pragma unroll
        for (int m = 0; m < 9; ++m) my_ms[m] = ms[m];
      }
    }
    if ((uint32_t)lane < nb)
      finish_normal(A, qstart + lane, my_cnt, my_ms, org);
    __syncwarp();
  }
}

// ---- the fast path: one warp per 2x2x2 group of cells, one query per lane ----------------------------------
// The points of a level-1 Morton parent (8 cells) are one contiguous range of the sorted cloud, and the union of
// their 27-cell neighbourhoods is the 4x4x4 block of cells around the parent.  The warp stages that block once
// (the group's own points first: they are the queries), then every lane serves ONE query: it walks the staged
// candidates (a shared-memory broadcast per candidate), keeps its K best (d^2, original index) keys in a sorted
// register list and sums the integer moments of everything inside the radius on the way.  No warp-wide
// selection, no re-reading: ~10 instructions per candidate and lane, an insertion only while the list still
// improves.  A superset of the 27 cells changes nothing in the exactness argument (rows with K-th d^2 <
// (cell+1)^2 are exact).  Left to the per-cell kernels: groups with more than GCAP candidates (dense scans) and
// -- hybrid set only -- groups where some query has more than max_nn points inside the radius (the max_nn
// nearest of them then need a selection).
constexpr int GCAP = 512;   // staged candidates per warp
constexpr int GKW = 4;      // warps per block

struct GroupScratch {
  uint32_t rs[64];   // start of the range in the sorted cloud, slot order = 2 * lane + {0, 1}
  uint32_t ro[65];   // its offset in the staging area (non-decreasing), ro[64] = M
  int32_t rows[32 * 16];
};

// one level of the cell hierarchy as the group kernel sees it
struct GridLevel {
  const uint32_t* cell_start;
  const uint64_t* hk;
  const uint32_t* hv;
  uint64_t hmask;
  uint32_t gmax[3];
  uint32_t guar2;  // (edge + 1)^2
  uint32_t unres_flag;  // FROM_MID for the finer table: the fallback then starts with the 27 cells of the coarse one
};
constexpr uint32_t FROM_MID = 0x80000000u;

template <int KT>
__device__ __forceinline__ void list_insert(unsigned long long (&L)[KT], unsigned long long key)
{
#pragma unroll
  for (int i = KT - 1; i > 0; --i) {
    const unsigned long long below = L[i - 1];
    L[i] = key < below ? below : (key < L[i] ? key : L[i]);
  }
  L[0] = key < L[0] ? key : L[0];
}

enum { GROUP_DONE = 0, GROUP_TOO_BIG = 1, GROUP_HYBRID_LEFT = 2 };

// staging order of the 4x4x4 block: the group's own 8 cells, then the 24 cells sharing a face with the core, the 24
// along its edges, the 8 corners -- near candidates first, so the lanes' K-th best tightens early and the far
// cells cause few insertions
__device__ const unsigned char block_order[64] = {
    21, 22, 25, 26, 37, 38, 41, 42, 5,  6,  9,  10, 17, 18, 20, 23, 24, 27, 29, 30, 33, 34, 36, 39, 40, 43, 45, 46, 53, 54, 57, 58,
    1,  2,  4,  7,  8,  11, 13, 14, 16, 19, 28, 31, 32, 35, 44, 47, 49, 50, 52, 55, 56, 59, 61, 62, 0,  3,  12, 15, 48, 51, 60, 63};

// Serves the queries [qstart, qstart + Q) -- the points of the 2x2x2 block of level-`G` cells whose parent has
// the Morton code P -- from the 4x4x4 block of cells around it.  Returns GROUP_TOO_BIG (nothing done) when the
// block holds more than GCAP points.
// `item` >= 0: the call comes from the main pass; a block that has to be streamed is not served but queued, one
// work item (item, batch) per 32 queries, so that a crowded group is shared by many warps.  Otherwise only the
// batches [b_begin, b_end) are served.
template <int KT>
__device__ int serve_group(const KnnArgs& A, const GridLevel& G, uint64_t P, uint32_t qstart, uint32_t Q, int4* sp,
                           GroupScratch* gs, int lane, bool allow_chunks, int64_t item, uint32_t b_begin, uint32_t b_end)
{
  // the 56 other cells of the 4x4x4 block
  const int64_t bx = 2 * (int64_t)morton_compact21(P) - 1, by = 2 * (int64_t)morton_compact21(P >> 1) - 1,
                bz = 2 * (int64_t)morton_compact21(P >> 2) - 1;
  uint32_t st2[2], ln2[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = block_order[2 * lane + h];  // staging slot 2 * lane + h
    const int dx = i & 3, dy = (i >> 2) & 3, dz = i >> 4;
    st2[h] = 0;
    ln2[h] = 0;
    const bool own = (dx == 1 || dx == 2) && (dy == 1 || dy == 2) && (dz == 1 || dz == 2);
    const int64_t nx = bx + dx, ny = by + dy, nz = bz + dz;
    if (!own && nx >= 0 && ny >= 0 && nz >= 0 && nx <= G.gmax[0] && ny <= G.gmax[1] && nz <= G.gmax[2]) {
      const uint32_t c = hash_lookup(G.hk, G.hv, G.hmask, morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
      if (c != 0xffffffffu) {
        st2[h] = __ldg(G.cell_start + c);
        ln2[h] = __ldg(G.cell_start + c + 1) - st2[h];
      }
    }
  }
  uint32_t incl = ln2[0] + ln2[1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= o) incl += v;
  }
  const uint32_t M = Q + __shfl_sync(FULL_MASK, incl, 31);
  const bool single = M <= (uint32_t)GCAP;  // else the block is streamed through the staging area in chunks
  if (!single && !allow_chunks)
    return GROUP_TOO_BIG;
  if (!single && item >= 0) {
    const uint32_t nbatch = (Q + 31) / 32;
    uint32_t slot = 0;
    if (lane == 0) slot = atomicAdd(A.n_dense, nbatch);
    slot = __shfl_sync(FULL_MASK, slot, 0);
    for (uint32_t b = lane; b < nbatch; b += 32) {
      A.dense_items[2 * (slot + b)] = (uint32_t)item;
      A.dense_items[2 * (slot + b) + 1] = b;
    }
    return GROUP_DONE;
  }
  __syncwarp();
  {
    const uint32_t base = Q + incl - (ln2[0] + ln2[1]);
    gs->rs[2 * lane] = st2[0];
    gs->ro[2 * lane] = base;
    gs->rs[2 * lane + 1] = st2[1];
    gs->ro[2 * lane + 1] = base + ln2[0];
    if (lane == 31) gs->ro[64] = M;
  }
  __syncwarp();
  // candidate e of the block: the group's own points (the queries) first, then the 56 ranges flattened
  auto stage = [&](uint32_t base, uint32_t m) {
    for (uint32_t t = lane; t < m; t += 32) {
      const uint32_t e = base + t;
      if (e < Q) {
        sp[t] = __ldg(A.pts + qstart + e);
      } else {
        int lo = 0, hi = 64;  // first slot with ro > e
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (gs->ro[mid] > e) hi = mid;
          else lo = mid + 1;
        }
        const int sidx = lo - 1;
        sp[t] = __ldg(A.pts + gs->rs[sidx] + (e - gs->ro[sidx]));
      }
    }
    __syncwarp();
  };
  // f(candidate) for every point of the block
  auto scan = [&](auto&& f) {
    if (single) {
      for (uint32_t j = 0; j < M; ++j) f(sp[j]);
    } else {
      for (uint32_t base = 0; base < M; base += GCAP) {
        const uint32_t m = min((uint32_t)GCAP, M - base);
        __syncwarp();
        stage(base, m);
        for (uint32_t j = 0; j < m; ++j) f(sp[j]);
      }
    }
  };
  if (single) stage(0, M);
  bool hybrid_left = false;
  for (uint32_t b0 = 32 * b_begin; b0 < Q && b0 < 32 * (uint64_t)b_end; b0 += 32) {
    const uint32_t nb = min(32u, Q - b0);
    const bool act = (uint32_t)lane < nb;
    const int4 q = __ldg(A.pts + qstart + (act ? b0 + lane : 0));
    unsigned long long L[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) L[i] = ~0ull;
    uint32_t cnt = 0;
    uint32_t a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    scan([&](const int4& c) {
      const int dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
      const uint32_t d2 = (uint32_t)(dx * dx) + (uint32_t)(dy * dy) + (uint32_t)(dz * dz);
      if (d2 < A.r2i) {
        ++cnt;
        a[0] += (uint32_t)dx; a[1] += (uint32_t)dy; a[2] += (uint32_t)dz;
        a[3] += (uint32_t)(dx * dx); a[4] += (uint32_t)(dx * dy); a[5] += (uint32_t)(dx * dz);
        a[6] += (uint32_t)(dy * dy); a[7] += (uint32_t)(dy * dz); a[8] += (uint32_t)(dz * dz);
      }
      const unsigned long long key = ((unsigned long long)d2 << 32) | (uint32_t)c.w;
      if (key < L[KT - 1]) list_insert<KT>(L, key);
    });
    // ---- rows: sorted positions of the K best, through shared memory for a coalesced store ----
    const uint32_t qpos = qstart + b0 + lane;
    __syncwarp();
    if (act) {
#pragma unroll
      for (int i = 0; i < KT; ++i)
        gs->rows[lane * KT + i] = L[i] == ~0ull ? -1 : (int32_t)__ldg(A.inv + (uint32_t)L[i]);
      const bool resolved = L[KT - 1] != ~0ull && (uint32_t)(L[KT - 1] >> 32) < G.guar2;
      if (!resolved) {
        const uint32_t slot = atomicAdd(A.n_unres, 1u);
        A.unres[slot] = qpos | G.unres_flag;
      }
    }
    __syncwarp();
    {
      int32_t* dst = A.nbr + (int64_t)(qstart + b0) * KT;
      for (uint32_t e = lane; e < nb * KT; e += 32) dst[e] = gs->rows[e];
    }
    // ---- hybrid set: everything inside the radius, or the max_nn nearest of it ----
    bool over = act && cnt > (uint32_t)A.max_nn;
    if (__any_sync(FULL_MASK, over)) {
      if (A.max_nn < KT) {
        hybrid_left = true;  // odd parameters: the per-cell kernels select
      } else {
        // T = the max_nn-th smallest key: the list holds the first KT, further passes fetch the next KT above
        // the last one known (keys are unique: the index is part of them)
        unsigned long long T = L[KT - 1];
        int more = over ? A.max_nn - KT : 0;
        while (__any_sync(FULL_MASK, more > 0)) {
          const unsigned long long lower = T;
#pragma unroll
          for (int i = 0; i < KT; ++i) L[i] = ~0ull;
          scan([&](const int4& c) {
            const int dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
            const uint32_t d2 = (uint32_t)(dx * dx) + (uint32_t)(dy * dy) + (uint32_t)(dz * dz);
            const unsigned long long key = ((unsigned long long)d2 << 32) | (uint32_t)c.w;
            if (more > 0 && key > lower && key < L[KT - 1]) list_insert<KT>(L, key);
          });
          if (more > 0) {
            const int take = more < KT ? more : KT;
            unsigned long long t = L[0];
#pragma unroll
            for (int i = 1; i < KT; ++i)
              if (i < take) t = L[i];
            T = t;
            more -= take;
          }
        }
        if (over) {
          cnt = (uint32_t)A.max_nn;
#pragma unroll
          for (int m = 0; m < 9; ++m) a[m] = 0;
        }
        scan([&](const int4& c) {
          const int dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
          const uint32_t d2 = (uint32_t)(dx * dx) + (uint32_t)(dy * dy) + (uint32_t)(dz * dz);
          const unsigned long long key = ((unsigned long long)d2 << 32) | (uint32_t)c.w;
          if (over && key <= T) {
            a[0] += (uint32_t)dx; a[1] += (uint32_t)dy; a[2] += (uint32_t)dz;
            a[3] += (uint32_t)(dx * dx); a[4] += (uint32_t)(dx * dy); a[5] += (uint32_t)(dx * dz);
            a[6] += (uint32_t)(dy * dy); a[7] += (uint32_t)(dy * dz); a[8] += (uint32_t)(dz * dz);
          }
        });
        over = false;
      }
    }
    if (act && !over) {
      int ms[9];
#pragma unroll
      for (int m = 0; m < 9; ++m) ms[m] = (int)a[m];
      finish_normal(A, qpos, (int)cnt, ms, make_int3(q.x, q.y, q.z));
    }
    __syncwarp();
  }
  return hybrid_left ? GROUP_HYBRID_LEFT : GROUP_DONE;
}

__device__ __forceinline__ void push_cells(const KnnArgs& A, uint32_t c0, int ng, uint32_t flag, int lane)
{
  uint32_t slot = 0;
  if (lane == 0) slot = atomicAdd(A.n_cell_list, (uint32_t)ng);
  slot = __shfl_sync(FULL_MASK, slot, 0);
  if (lane < ng) A.cell_list[slot + lane] = (c0 + lane) | flag;
}

template <int KT>
__global__ void __launch_bounds__(GKW * 32) knn_groups_kernel(KnnArgs A, GridLevel G0, GridLevel G1, int have_mid)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int4* sp = reinterpret_cast<int4*>(smem) + (size_t)w * GCAP;
  GroupScratch* gs = reinterpret_cast<GroupScratch*>(smem + (size_t)GKW * GCAP * 16) + w;
  const uint32_t n_chunks = (A.n_cells + 31) / 32;
  for (;;) {
    uint32_t chunk = 0;
    if (lane == 0) chunk = atomicAdd(A.next_chunk, 1u);
    chunk = __shfl_sync(FULL_MASK, chunk, 0);
    if (chunk >= n_chunks)
      break;
    // group leaders among the chunk's 32 cells: first cell of its level-1 parent
    const uint32_t myc = chunk * 32 + lane;
    uint64_t mykey = 0;
    bool leader = false;
    if (myc < A.n_cells) {
      mykey = __ldg(A.cell_key + myc);
      leader = myc == 0 || (__ldg(A.cell_key + myc - 1) >> 3) != (mykey >> 3);
    }
    uint32_t leaders = __ballot_sync(FULL_MASK, leader);
    while (leaders) {
      const int ll = __ffs(leaders) - 1;
      leaders &= leaders - 1;
      const uint32_t c0 = chunk * 32 + ll;
      const uint64_t P = __shfl_sync(FULL_MASK, mykey, ll) >> 3;
      // cells of the group (contiguous: the keys are sorted)
      bool in = false;
      if (lane < 8 && c0 + lane < A.n_cells) in = (__ldg(A.cell_key + c0 + lane) >> 3) == P;
      const int ng = __popc(__ballot_sync(FULL_MASK, in));
      const uint32_t qstart = __ldg(A.cell_start + c0);
      const uint32_t Q = __ldg(A.cell_start + c0 + ng) - qstart;
      const int r = serve_group<KT>(A, G0, P, qstart, Q, sp, gs, lane, !have_mid, (int64_t)c0, 0u, 0xffffffffu);
      if (r == GROUP_HYBRID_LEFT) {
        push_cells(A, c0, ng, HYB_ONLY, lane);
      } else if (r == GROUP_TOO_BIG) {
        if (!have_mid) {
          push_cells(A, c0, ng, 0u, lane);  // dense: the per-cell kernels take the group's cells
        } else {
          // dense: every cell of the group becomes a group of the finer level (its 8 half-size cells)
          for (int i = 0; i < ng; ++i) {
            const uint32_t cs = __ldg(A.cell_start + c0 + i);
            const uint32_t cq = __ldg(A.cell_start + c0 + i + 1) - cs;
            const int r1 = serve_group<KT>(A, G1, __ldg(A.cell_key + c0 + i), cs, cq, sp, gs, lane, true,
                                           (int64_t)((c0 + i) | 0x80000000u), 0u, 0xffffffffu);
            if (r1 != GROUP_DONE) push_cells(A, c0 + i, 1, r1 == GROUP_HYBRID_LEFT ? HYB_ONLY : 0u, lane);
          }
        }
      }
      __syncwarp();
    }
  }
}

// one warp per queued batch of a crowded group: the block is streamed through the staging area
template <int KT>
__global__ void __launch_bounds__(GKW * 32) knn_dense_kernel(KnnArgs A, GridLevel G0, GridLevel G1)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int4* sp = reinterpret_cast<int4*>(smem) + (size_t)w * GCAP;
  GroupScratch* gs = reinterpret_cast<GroupScratch*>(smem + (size_t)GKW * GCAP * 16) + w;
  const uint32_t nitems = *A.n_dense;
  for (;;) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(A.next_dense, 1u);
    it = __shfl_sync(FULL_MASK, it, 0);
    if (it >= nitems)
      break;
    const uint32_t item = A.dense_items[2 * it], batch = A.dense_items[2 * it + 1];
    const uint32_t c = item & 0x7fffffffu;
    int r;
    int ng = 1;
    if (item & 0x80000000u) {  // a cell of the coarse table as a group of the finer one
      const uint32_t cs = __ldg(A.cell_start + c);
      const uint32_t cq = __ldg(A.cell_start + c + 1) - cs;
      r = serve_group<KT>(A, G1, __ldg(A.cell_key + c), cs, cq, sp, gs, lane, true, -1, batch, batch + 1);
    } else {
      const uint64_t P = __ldg(A.cell_key + c) >> 3;
      bool in = false;
      if (lane < 8 && c + lane < A.n_cells) in = (__ldg(A.cell_key + c + lane) >> 3) == P;
      ng = __popc(__ballot_sync(FULL_MASK, in));
      const uint32_t qstart = __ldg(A.cell_start + c);
      const uint32_t Q = __ldg(A.cell_start + c + ng) - qstart;
      r = serve_group<KT>(A, G0, P, qstart, Q, sp, gs, lane, true, -1, batch, batch + 1);
    }
    // odd parameters (max_nn < K): the per-cell kernels select the hybrid sets of the whole group
    (void)r;
    if (batch == 0 && A.max_nn < KT) push_cells(A, c, ng, HYB_ONLY, lane);
    __syncwarp();
  }
}

// ---- ring-expansion fallback: one warp per unresolved query, lane l holds the l-th best ------------
__device__ __forceinline__ bool cand_less(uint64_t d2a, uint32_t ia, uint64_t d2b, uint32_t ib)
{
  return d2a < d2b || (d2a == d2b && ia < ib);
}

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* __restrict__ a, uint32_t n, uint64_t v)
{
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(a + mid) < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(KTHREADS) knn_fallback_kernel(KnnArgs A, int max_level)
{
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t nun = *A.n_unres;
  const uint32_t kmask = A.K >= 32 ? 0xffffffffu : ((1u << A.K) - 1u);
  for (uint32_t i = blockIdx.x * KW + w; i < nun; i += gridDim.x * KW) {
    const uint32_t entry = A.unres[i];
    const uint32_t qpos = entry & ~FROM_MID;
    const int4 q = __ldg(A.pts + qpos);
    const uint32_t qcx = (uint32_t)q.x / (uint32_t)A.cell, qcy = (uint32_t)q.y / (uint32_t)A.cell,
                   qcz = (uint32_t)q.z / (uint32_t)A.cell;
    uint64_t bd2 = ~0ull;
    uint32_t bidx = 0xffffffffu, bpos = 0xffffffffu;
    for (int L = (entry & FROM_MID) ? 0 : 1; L <= max_level; ++L) {
      bd2 = ~0ull; bidx = 0xffffffffu; bpos = 0xffffffffu;
      const uint32_t X = qcx >> L, Y = qcy >> L, Z = qcz >> L;
      const uint32_t gx = A.gmax[0] >> L, gy = A.gmax[1] >> L, gz = A.gmax[2] >> L;
      uint32_t rs = 0, re = 0;
      if (lane < 27) {
        int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
        int64_t nx = (int64_t)X + dx, ny = (int64_t)Y + dy, nz = (int64_t)Z + dz;
        if (nx >= 0 && ny >= 0 && nz >= 0 && nx <= gx && ny <= gy && nz <= gz) {
          uint64_t P = morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
          rs = lower_bound_u64(A.keys, (uint32_t)A.n, P << (3 * L));
          re = lower_bound_u64(A.keys, (uint32_t)A.n, (P + 1) << (3 * L));
        }
      }
      for (int c = 0; c < 27; ++c) {
        const uint32_t s = __shfl_sync(FULL_MASK, rs, c), e = __shfl_sync(FULL_MASK, re, c);
        for (uint32_t t0 = s; t0 < e; t0 += 32) {
          const uint32_t t = t0 + lane;
          uint64_t d2 = ~0ull;
          uint32_t idx = 0xffffffffu;
          if (t < e) {
            int4 p = __ldg(A.pts + t);
            int64_t dx = (int64_t)p.x - q.x, dy = (int64_t)p.y - q.y, dz = (int64_t)p.z - q.z;
            d2 = (uint64_t)(dx * dx + dy * dy + dz * dz);
            idx = (uint32_t)p.w;
          }
          uint64_t kd2 = __shfl_sync(FULL_MASK, bd2, A.K - 1);
          uint32_t kidx = __shfl_sync(FULL_MASK, bidx, A.K - 1);
          uint32_t mask = __ballot_sync(FULL_MASK, t < e && cand_less(d2, idx, kd2, kidx));
          while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint64_t cd2 = __shfl_sync(FULL_MASK, d2, src);
            const uint32_t cidx = __shfl_sync(FULL_MASK, idx, src);
            const uint32_t cpos = t0 + src;
            // position = number of list entries ahead of the candidate
            const uint32_t ahead = __ballot_sync(FULL_MASK, cand_less(bd2, bidx, cd2, cidx)) & kmask;
            const int pos = __popc(ahead);
            const uint64_t ud2 = __shfl_up_sync(FULL_MASK, bd2, 1);
            const uint32_t uidx = __shfl_up_sync(FULL_MASK, bidx, 1);
            const uint32_t upos = __shfl_up_sync(FULL_MASK, bpos, 1);
            if (pos < A.K) {
              if (lane > pos) { bd2 = ud2; bidx = uidx; bpos = upos; }
              else if (lane == pos) { bd2 = cd2; bidx = cidx; bpos = cpos; }
            }
          }
        }
      }
      const uint64_t kd2 = __shfl_sync(FULL_MASK, bd2, A.K - 1);
      const uint64_t reach = ((uint64_t)A.cell << L) + 1;
      const bool covered = (X <= 1) && (Y <= 1) && (Z <= 1) && (X + 1 >= gx) && (Y + 1 >= gy) && (Z + 1 >= gz);
      if (covered || (kd2 != ~0ull && kd2 < reach * reach))
        break;
    }
    if (lane < A.K)
      A.nbr[(int64_t)qpos * A.K + lane] = (bd2 == ~0ull) ? -1 : (int32_t)bpos;
    __syncwarp();
  }
}

// ---- exports ---------------------------------------------------------------------------------------------
__global__ void export_rows_kernel(const int32_t* __restrict__ nbr, const int4* __restrict__ pts, int64_t n, int K,
                                   int32_t* __restrict__ out)
{
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * K)
    return;
  int64_t s = e / K;
  int k = (int)(e - s * K);
  int32_t v = nbr[e];
  int32_t o = pts[s].w;
  out[(int64_t)o * K + k] = v < 0 ? -1 : pts[v].w;
}

__global__ void export_normals_kernel(const double* __restrict__ nrm, const double* __restrict__ curv,
                                      const int4* __restrict__ pts, int64_t n, double* __restrict__ out_n,
                                      double* __restrict__ out_c)
{
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
    return;
  int64_t o = pts[s].w;
  if (out_n) {
    out_n[3 * o] = nrm[3 * s];
    out_n[3 * o + 1] = nrm[3 * s + 1];
    out_n[3 * o + 2] = nrm[3 * s + 2];
  }
  if (out_c)
    out_c[o] = curv[s];
}

__global__ void import_rows_kernel(const int32_t* __restrict__ in, const uint32_t* __restrict__ inv, int64_t n, int K,
                                   int32_t* __restrict__ nbr)
{
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * K)
    return;
  int64_t o = e / K;
  int k = (int)(e - o * K);
  int32_t v = in[e];
  nbr[(int64_t)inv[o] * K + k] = (v < 0 || v >= n) ? -1 : (int32_t)inv[v];
}

__global__ void import_normals_kernel(const double* __restrict__ in, const uint32_t* __restrict__ inv, int64_t n,
                                      double* __restrict__ nrm, unsigned long long* __restrict__ n_nonunit)
{
  int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n)
    return;
  int64_t s = inv[o];
  const double a = in[3 * o], b = in[3 * o + 1], cc = in[3 * o + 2];
  nrm[3 * s] = a;
  nrm[3 * s + 1] = b;
  nrm[3 * s + 2] = cc;
  // the grower's approximate-model margins (grow.cuh geo_test_margin) assume |n_i| ~ 1: count the others
  const double q = a * a + b * b + cc * cc;
  if (!(q > 0.999 && q < 1.001)) atomicAdd(n_nonunit, 1ull);
}

}  // namespace

int stage_knn(bseg_ctx* c, const bseg_params* p)
{
  const int64_t n = c->n;
  c->normals_nonunit = false;
  if (n == 0)
    return 0;
  RC_CHECK(dev_ensure(c, c->nbr, (size_t)n * p->K * 4));
  RC_CHECK(dev_ensure(c, c->nrm, (size_t)n * 24));
  RC_CHECK(dev_ensure(c, c->curv, (size_t)n * 8));
  const size_t max_items = (size_t)c->n_cells + (size_t)n / 32 + 8;
  RC_CHECK(dev_ensure(c, c->worklist, ((size_t)n + 4 * max_items + (size_t)c->n_cells + 16) * 4));
  RC_CHECK(dev_ensure(c, c->counters, 192 * sizeof(uint64_t)));

  KnnArgs A;
  A.pts = dptr<int4>(c->pts);
  A.keys = dptr<uint64_t>(c->keys[c->sort_sel]);
  A.cell_start = dptr<uint32_t>(c->cell_start);
  A.cell_key = dptr<uint64_t>(c->cell_key);
  A.n_cells = (uint32_t)c->n_cells;
  A.hk = dptr<uint64_t>(c->hash_keys);
  A.hv = dptr<uint32_t>(c->hash_vals);
  A.hmask = c->hash_mask;
  A.cell = c->cell;
  for (int k = 0; k < 3; ++k)
    A.gmax[k] = (uint32_t)(c->mx[k] - c->mn[k]) / (uint32_t)c->cell;
  A.K = p->K;
  A.max_nn = p->max_nn;
  {
    double r2 = p->radius * p->radius;
    uint32_t r2i = (uint32_t)r2;
    if ((double)r2i < r2) ++r2i;  // d2 < r2  <=>  d2 < ceil(r2) for integer d2
    A.r2i = r2i;
  }
  A.guar2 = (uint32_t)(c->cell + 1) * (uint32_t)(c->cell + 1);
  A.nbr = dptr<int32_t>(c->nbr);
  A.nrm = dptr<double>(c->nrm);
  A.curv = dptr<double>(c->curv);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(dptr<uint64_t>(c->counters)) + 32;
  A.n_big = cnt;
  A.n_unres = cnt + 1;
  A.big_cells = dptr<uint32_t>(c->worklist);
  A.unres = dptr<uint32_t>(c->worklist) + 2 * max_items;
  A.cell_list = A.unres + n;
  A.n_cell_list = cnt + 2;
  A.next_chunk = cnt + 3;
  A.n_dense = cnt + 4;
  A.next_dense = cnt + 5;
  A.dense_items = A.cell_list + c->n_cells;
  A.inv = dptr<uint32_t>(c->inv);
  A.n = n;

  STAGE_BEGIN(c, EV_KNN);
  CU_CHECK(c, cudaMemsetAsync(cnt, 0, 6 * sizeof(uint32_t), c->stream));
  const size_t smem = (size_t)KW * CAP * 20 + KW * sizeof(WarpScratch);
  // the group kernel serves K = 15 (the reference) and 16; BSEG_KNN_GROUPS=0 forces the per-cell kernels
  const bool groups_on = !(getenv("BSEG_KNN_GROUPS") && atoi(getenv("BSEG_KNN_GROUPS")) == 0);
  const bool groups = groups_on && (p->K == 15 || p->K == 16);
  if (groups) {
    const size_t gsmem = (size_t)GKW * GCAP * 16 + GKW * sizeof(GroupScratch);
    if (!c->attr_knn_set) {
      CU_CHECK(c, cudaFuncSetAttribute(knn_groups_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
      CU_CHECK(c, cudaFuncSetAttribute(knn_groups_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
      CU_CHECK(c, cudaFuncSetAttribute(knn_dense_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
      CU_CHECK(c, cudaFuncSetAttribute(knn_dense_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
      c->attr_knn_set = true;
    }
    GridLevel G0, G1;
    G0.cell_start = A.cell_start; G0.hk = A.hk; G0.hv = A.hv; G0.hmask = A.hmask; G0.guar2 = A.guar2;
    G0.unres_flag = 0;
    for (int k = 0; k < 3; ++k) G0.gmax[k] = A.gmax[k];
    G1 = G0;
    const bool mid = c->have_mid && (c->cell % 2) == 0 && (double)(c->cell / 2) >= p->radius;
    if (mid) {
      const int32_t cell2 = c->cell / 2;
      G1.cell_start = dptr<uint32_t>(c->cell_start2);
      G1.hk = dptr<uint64_t>(c->hash_keys2);
      G1.hv = dptr<uint32_t>(c->hash_vals2);
      G1.hmask = c->hash_mask2;
      G1.guar2 = (uint32_t)(cell2 + 1) * (uint32_t)(cell2 + 1);
      G1.unres_flag = FROM_MID;
      for (int k = 0; k < 3; ++k) G1.gmax[k] = (uint32_t)(c->mx[k] - c->mn[k]) / (uint32_t)cell2;
    }
    const unsigned gb = (unsigned)(c->num_sms * 6);
    if (p->K == 15) knn_groups_kernel<15><<<gb, GKW * 32, gsmem, c->stream>>>(A, G0, G1, mid ? 1 : 0);
    else knn_groups_kernel<16><<<gb, GKW * 32, gsmem, c->stream>>>(A, G0, G1, mid ? 1 : 0);
    KLAUNCH_CHECK(c);
    if (p->K == 15) knn_dense_kernel<15><<<gb, GKW * 32, gsmem, c->stream>>>(A, G0, G1);
    else knn_dense_kernel<16><<<gb, GKW * 32, gsmem, c->stream>>>(A, G0, G1);
    KLAUNCH_CHECK(c);
    knn_cells_kernel<<<c->num_sms * 4, KTHREADS, smem, c->stream>>>(A, 1);
  } else {
    knn_cells_kernel<<<c->num_sms * 8, KTHREADS, smem, c->stream>>>(A, 0);
  }
  KLAUNCH_CHECK(c);
  knn_big_cells_kernel<<<c->num_sms * 4, KTHREADS, smem, c->stream>>>(A);
  KLAUNCH_CHECK(c);
  STAGE_END(c, EV_KNN);

  STAGE_BEGIN(c, EV_KNN_FB);
  int max_level = 1;
  {
    uint32_t g = A.gmax[0] > A.gmax[1] ? A.gmax[0] : A.gmax[1];
    if (A.gmax[2] > g) g = A.gmax[2];
    while ((g >> max_level) > 0) ++max_level;
    ++max_level;
  }
  knn_fallback_kernel<<<c->num_sms * 8, KTHREADS, 0, c->stream>>>(A, max_level);
  KLAUNCH_CHECK(c);
  STAGE_END(c, EV_KNN_FB);
  uint32_t h[2];
  RC_CHECK(read_back(c, h, cnt, sizeof(h)));
  c->tm.n_unresolved = h[1];
  c->tm.n_big_cells = h[0];
  return 0;
}

// ---- halo sufficiency of a slab (include/bseg.h: bseg_halo_check) ------------------------------------------------
namespace {
__global__ void halo_check_kernel(const int4* __restrict__ pts, const int32_t* __restrict__ nbr, int K, int64_t n,
                                  int64_t n_owned, int32_t x_lo, int32_t x_hi, int32_t halo,
                                  unsigned long long* __restrict__ count)
{
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
    return;
  const int4 q = __ldg(pts + s);
  if (q.w >= n_owned)
    return;  // declared a halo copy (bseg_set_owned): only a neighbour
  // a face without a neighbour rank (outer face of the tile) is INT32_MIN / INT32_MAX: nothing lies beyond it
  const bool has_l = x_lo != INT32_MIN, has_r = x_hi != INT32_MAX;
  if ((has_l && q.x < x_lo) || (has_r && q.x >= x_hi))
    return;  // outside the slab: a halo copy by position
  // the points this rank does NOT hold lie more than `halo` beyond a face: at least reach = halo + (distance of q
  // to that face) away from q
  int64_t reach = INT64_MAX;
  if (has_l) reach = (int64_t)q.x - x_lo + halo;
  if (has_r) {
    const int64_t r = (int64_t)x_hi - 1 - q.x + halo;
    reach = r < reach ? r : reach;
  }
  if (reach == INT64_MAX)
    return;
  int32_t last = -1;
  for (int j = K - 1; j >= 0 && last < 0; --j) last = __ldg(nbr + s * K + j);
  bool bad = last < 0;  // fewer than K points in reach at all
  if (!bad) {
    bad = __ldg(nbr + s * K + (K - 1)) < 0;  // a short row: the missing entries may lie beyond the halo
    const int4 r = __ldg(pts + last);
    const int64_t dx = (int64_t)r.x - q.x, dy = (int64_t)r.y - q.y, dz = (int64_t)r.z - q.z;
    bad = bad || dx * dx + dy * dy + dz * dz > reach * reach;
  }
  if (bad) atomicAdd(count, 1ull);
}
}  // namespace

int stage_halo_check(bseg_ctx* c, int32_t x_lo, int32_t x_hi, int32_t halo, int64_t* n_unresolved)
{
  *n_unresolved = 0;
  if (c->n == 0)
    return 0;
  RC_CHECK(dev_ensure(c, c->counters, 192 * sizeof(uint64_t)));
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(dptr<uint64_t>(c->counters)) + 190;
  CU_CHECK(c, cudaMemsetAsync(cnt, 0, sizeof(*cnt), c->stream));
  halo_check_kernel<<<(unsigned)ceil_div64(c->n, 256), 256, 0, c->stream>>>(dptr<int4>(c->pts), dptr<int32_t>(c->nbr), c->K,
                                                                          c->n, c->n_owned, x_lo, x_hi, halo, cnt);
  KLAUNCH_CHECK(c);
  unsigned long long h = 0;
  RC_CHECK(read_back(c, &h, cnt, sizeof(h)));
  *n_unresolved = (int64_t)h;
  return 0;
}

int stage_export_knn(bseg_ctx* c, const bseg_params* p, int32_t* h_neigh, double* h_normals, double* h_curv)
{
  const int64_t n = c->n;
  if (n == 0)
    return 0;
  const int K = p->K;
  STAGE_BEGIN(c, EV_D2H);
  if (h_neigh) {
    RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)n * K * 4));
    export_rows_kernel<<<(unsigned)ceil_div64(n * K, 256), 256, 0, c->stream>>>(dptr<int32_t>(c->nbr), dptr<int4>(c->pts),
                                                                              n, K, dptr<int32_t>(c->out_tmp));
    KLAUNCH_CHECK(c);
    CU_CHECK(c, cudaMemcpyAsync(h_neigh, c->out_tmp.p, (size_t)n * K * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  if (h_normals || h_curv) {
    RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)n * 32));
    double* on = dptr<double>(c->out_tmp);
    double* oc = on + 3 * n;
    export_normals_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(
        dptr<double>(c->nrm), dptr<double>(c->curv), dptr<int4>(c->pts), n, h_normals ? on : nullptr,
        h_curv ? oc : nullptr);
    KLAUNCH_CHECK(c);
    if (h_normals)
      CU_CHECK(c, cudaMemcpyAsync(h_normals, on, (size_t)n * 24, cudaMemcpyDeviceToHost, c->stream));
    if (h_curv)
      CU_CHECK(c, cudaMemcpyAsync(h_curv, oc, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  STAGE_END(c, EV_D2H);
  return 0;
}

int stage_override(bseg_ctx* c, const bseg_params* p, const int32_t* neigh, const double* normals, bool on_device)
{
  const int64_t n = c->n;
  if (n == 0)
    return 0;
  const int K = p->K;
  if (neigh) {
    const int32_t* src = neigh;
    if (!on_device) {
      RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)n * K * 4));
      CU_CHECK(c, cudaMemcpyAsync(c->out_tmp.p, neigh, (size_t)n * K * 4, cudaMemcpyHostToDevice, c->stream));
      src = dptr<int32_t>(c->out_tmp);
    }
    import_rows_kernel<<<(unsigned)ceil_div64(n * K, 256), 256, 0, c->stream>>>(src, dptr<uint32_t>(c->inv), n, K,
                                                                              dptr<int32_t>(c->nbr));
    KLAUNCH_CHECK(c);
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  if (normals) {
    const double* src = normals;
    if (!on_device) {
      RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)n * 24));
      CU_CHECK(c, cudaMemcpyAsync(c->out_tmp.p, normals, (size_t)n * 24, cudaMemcpyHostToDevice, c->stream));
      src = dptr<double>(c->out_tmp);
    }
    RC_CHECK(dev_ensure(c, c->counters, 192 * sizeof(uint64_t)));
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(dptr<uint64_t>(c->counters)) + 189;
    CU_CHECK(c, cudaMemsetAsync(cnt, 0, sizeof(*cnt), c->stream));
    import_normals_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(src, dptr<uint32_t>(c->inv), n,
                                                                             dptr<double>(c->nrm), cnt);
    KLAUNCH_CHECK(c);
    unsigned long long h = 0;
    RC_CHECK(read_back(c, &h, cnt, sizeof(h)));
    c->normals_nonunit = h != 0;  // un-normalised caller normals: the grower decides with the exact model only
  }
  return 0;
}

// rows / normals of the last kNN stage in ORIGINAL index space, left on the device (multi-GPU: the rows of a slab
// travel to the rank that grows the tile without a host round trip)
int stage_export_knn_device(bseg_ctx* c, const bseg_params* p, const int32_t** d_neigh, const double** d_normals)
{
  const int64_t n = c->n;
  const int K = p->K;
  RC_CHECK(dev_ensure(c, c->x_neigh, (size_t)n * K * 4));
  RC_CHECK(dev_ensure(c, c->x_normals, (size_t)n * 24));
  if (n > 0) {
    export_rows_kernel<<<(unsigned)ceil_div64(n * K, 256), 256, 0, c->stream>>>(dptr<int32_t>(c->nbr), dptr<int4>(c->pts),
                                                                              n, K, dptr<int32_t>(c->x_neigh));
    KLAUNCH_CHECK(c);
    export_normals_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(dptr<double>(c->nrm), dptr<double>(c->curv),
                                                                             dptr<int4>(c->pts), n,
                                                                             dptr<double>(c->x_normals), nullptr);
    KLAUNCH_CHECK(c);
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  if (d_neigh) *d_neigh = dptr<int32_t>(c->x_neigh);
  if (d_normals) *d_normals = dptr<double>(c->x_normals);
  return 0;
}

// binning only + room for rows / normals that arrive from elsewhere (bseg_import_neigh_normals_device)
int stage_alloc_knn_outputs(bseg_ctx* c, const bseg_params* p)
{
  const int64_t n = c->n;
  RC_CHECK(dev_ensure(c, c->nbr, (size_t)n * p->K * 4));
  RC_CHECK(dev_ensure(c, c->nrm, (size_t)n * 24));
  RC_CHECK(dev_ensure(c, c->curv, (size_t)n * 8));
  if (n > 0)
    CU_CHECK(c, cudaMemsetAsync(c->curv.p, 0, (size_t)n * 8, c->stream));
  return 0;
}
