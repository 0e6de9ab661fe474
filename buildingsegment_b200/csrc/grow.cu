// grow.cu -- north_star stage (4): plane region growing, order-faithful.
//
// Replaces seg_plane::get_planes / Broad / set_plane_color (my_function.cpp:180-275).  The
// reference is a sequential, index-ordered, seeded greedy grower whose result depends on the
// point order, so "label-exact" means reproducing its sequential
// semantics, not connected components.  What is reproduced, line by line:
//   :184-191  seeds are visited in original index order; a seed is appended but NOT marked (Q1)
//   :224-236  neighbours 1..K-1 of the node are tested against the model as it was BEFORE the call:
//             planeIdx <= 0, |(p - centre) . n| <= th_thickness, n . n_i >= th_dot; accepted
//             points are marked and appended in neighbour order
//   :238-239  at depth 0 all K-1 must be accepted, else the call fails and the marks stay (Q2)
//   :241-250  model = normalised running sum of normals (fp64, acceptance order) and the int32
//             wrapping centroid sum divided through uint64 (Q6); running sums are bit-identical
//             to the reference's from-scratch re-summation (Q10)
//   :252-255  depth-first descent into the accepted points, in order
//   :199-209  commit when pointIdx.size() > th_point_count, else un-mark the list
//
// Engines
//   sequential (grow_mode 1, and the head runner of mode 0): ONE warp walks the seeds in order;
//     lane j tests neighbour j, the model lives in registers, the DFS frames in a global stack.
//     O(1) per Broad call instead of the reference's O(P) re-summation.
//   speculative (grow_mode 0): transactions (one per free seed) of a window of indices run
//     concurrently against the committed state, reserving every point they accept with
//     atomicMin(original seed index).  A transaction's decisions depend only on the state of the
//     points it accepted, so it is valid iff it still holds all its reservations and its seed was
//     not accepted by a lower transaction; the valid PREFIX of the window commits, the first
//     invalid seed is re-run at the head.  Deterministic: the result is the sequential one.
//
// Layout: everything is indexed by Morton-sorted position s (neighbours are close in memory);
// seed order and every exported index use the original numbering (pts[s].w, inv[orig]).
// state[s] = -1 free, else ORIGINAL index of the seed whose transaction marked it; plane ids are
// assigned at the end: id(owner) = 1 + #committed plane seeds below owner (cur_planeId at that time).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "grow.cuh"

namespace {

// ---- sequential engine -----------------------------------------------------------------------------
// One warp.  Runs transactions in seed order starting at ctl[FRONTIER] until `max_tx` transactions
// or `max_steps` Broad calls were executed (checked between transactions) or the cloud is done.
// NOTIFY (head runner of the speculative engine): dooms in-flight transactions it overlaps, stops at
// seeds that own a grower slot, and -- unless allow_growers -- at any depth-0 success, which is left
// to a slot so that it grows concurrently with the others.
template <bool NOTIFY>
__global__ void __launch_bounds__(32) grow_seq_kernel(GrowArgs A, unsigned long long max_tx, unsigned long long max_steps,
                                                     int allow_growers)
{
  const int lane = threadIdx.x;
  int64_t frontier = (int64_t)A.ctl[CTL_FRONTIER];
  int64_t pool_used = (int64_t)A.ctl[CTL_POOL];
  int64_t n_planes = (int64_t)A.ctl[CTL_PLANES];
  unsigned long long steps = 0, ntx = 0;
  bool stop = false;

  for (int64_t base = frontier; base < A.n && !stop; base += 32) {
    const int64_t i_l = base + lane;
    uint32_t s_l = 0;
    bool free_l = false;
    if (i_l < A.n) {
      s_l = __ldg(A.inv + i_l);
      free_l = __ldcg(A.state + 2 * (int64_t)(s_l)) == -1;
    }
    uint32_t todo = __ballot_sync(FULL_MASK, free_l);
    frontier = base;
    while (todo) {
      const int b = __ffs(todo) - 1;
      if (ntx >= max_tx || steps >= max_steps) {
        frontier = base + b;
        stop = true;
        break;
      }
      todo &= todo - 1;
      const int64_t seed_i = base + b;
      const uint32_t seed_s = __shfl_sync(FULL_MASK, s_l, b);
      if (__ldcg(A.state + 2 * (int64_t)(seed_s)) != -1)
        continue;  // taken by a transaction of this very batch
      if (NOTIFY && A.slotof[seed_i] >= 0) {  // a speculative grower owns this seed: leave it to the window
        frontier = seed_i;
        stop = true;
        break;
      }
      if (pool_used + A.n + 2 > A.pool_cap || n_planes >= A.planes_cap) {
        if (lane == 0) A.ctl[CTL_ERR] = 1;
        frontier = seed_i;
        stop = true;
        break;
      }
      FlatStore st;
      st.list = A.pool + pool_used;
      st.stack = A.stack;
      if (lane == 0) st.list[0] = (int32_t)seed_s;
      TxState t;
      tx_begin(t, A, seed_s);
      const TxOutcome out = tx_run<NOTIFY ? MODE_SEQ_NOTIFY : MODE_SEQ, 0>(A, st, t, seed_i, ~0ull,
                                                                         NOTIFY && !allow_growers, lane, steps);
      if (out == TX_IS_GROWER) {
        frontier = seed_i;
        stop = true;
        break;
      }
      ++ntx;
      if (out == TX_FAILED)
        continue;  // :238-239 -- the marks stay (orphans)
      if (t.len > A.th_count) {  // :199
        if (lane == 0) {
          PlaneRec r;
          r.seed = (int32_t)seed_i; r.pad = 0;
          r.off = pool_used; r.len = t.len;
          r.nrm[0] = t.m.mn0; r.nrm[1] = t.m.mn1; r.nrm[2] = t.m.mn2;
          r.ctr[0] = t.m.mc0; r.ctr[1] = t.m.mc1; r.ctr[2] = t.m.mc2; r.pad2 = 0;
          A.planes[n_planes] = r;
        }
        pool_used += t.len;
        ++n_planes;
      } else {
        for (int64_t e = lane; e < t.len; e += 32) {  // :203-209
          A.state[2 * (int64_t)st.list[e]] = -1;
          if (NOTIFY) atomicCAS(A.res + 2 * (int64_t)(st.list[e]), (uint32_t)seed_i, RES_FREE);
        }
        __syncwarp();
      }
    }
    if (!stop)
      frontier = base + 32 < A.n ? base + 32 : A.n;
  }
  if (lane == 0) {
    A.ctl[CTL_FRONTIER] = (unsigned long long)frontier;
    A.ctl[CTL_POOL] = (unsigned long long)pool_used;
    A.ctl[CTL_PLANES] = (unsigned long long)n_planes;
    A.ctl[CTL_STEPS] += steps;
    A.ctl[CTL_TX] += ntx;
  }
}

// rowdup[s] = 1 when columns 1..K-1 of row s name some point twice (only then the per-step dedupe is needed)
__global__ void rowdup_kernel(const int32_t* __restrict__ nbr, int K, int64_t n, uint8_t* __restrict__ rowdup)
{
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
    return;
  const int32_t* row = nbr + s * K;
  bool dup = false;
  for (int j = 2; j < K; ++j) {
    const int32_t id = __ldg(row + j);
    if (id < 0)
      continue;
    for (int j2 = 1; j2 < j; ++j2) dup |= __ldg(row + j2) == id;
  }
  rowdup[s] = dup ? 1 : 0;
}

// the grower's int32 arithmetic in the CALLER's coordinates (bseg_set_grow_offset): pts + offset, wrapping as int32 does
__global__ void offset_pts_kernel(const int4* __restrict__ pts, int64_t n, int32_t o0, int32_t o1, int32_t o2,
                                  int4* __restrict__ out)
{
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n)
    return;
  int4 p = pts[s];
  p.x = (int32_t)((uint32_t)p.x + (uint32_t)o0);
  p.y = (int32_t)((uint32_t)p.y + (uint32_t)o1);
  p.z = (int32_t)((uint32_t)p.z + (uint32_t)o2);
  out[s] = p;
}

// "radius-search growing" (bseg_params.grow_radius): the grower's list of a point = the entries of its K-row with
// d^2 < r^2 (column 0 is kept whatever it is: Broad skips it blindly, my_function.cpp:224); the rest do not exist
__global__ void mask_rows_kernel(const int32_t* __restrict__ nbr, const int4* __restrict__ pts, int64_t n, int K,
                                 unsigned long long r2, int32_t* __restrict__ out)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * K)
    return;
  const int64_t s = e / K;
  const int j = (int)(e - s * K);
  int32_t id = nbr[e];
  if (j > 0 && id >= 0) {
    const int4 a = __ldg(pts + s), b = __ldg(pts + id);
    const long long dx = (long long)a.x - b.x, dy = (long long)a.y - b.y, dz = (long long)a.z - b.z;
    if ((unsigned long long)(dx * dx + dy * dy + dz * dz) >= r2) id = -1;
  }
  out[e] = id;
}

// ---- finalize: plane ids, planeIdx and label in original order ------------------------------------------
// id(owner) = 1 + #plane seeds < owner; label = id when the owner IS a plane seed, else the id of the
// plane seeded at the point itself (a seed is in its own pointIdx without being marked), else 0.
__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int n, int32_t v)
{
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__global__ void finalize_kernel(const int32_t* __restrict__ state, const uint32_t* __restrict__ inv,
                                const int32_t* __restrict__ seeds, int n_planes, int64_t n,
                                int32_t* __restrict__ pidx, int32_t* __restrict__ label)
{
  int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n)
    return;
  int32_t owner = state[2 * (int64_t)inv[o]];  // (owner, reservation) records
  int32_t pi = -1, lb = 0;
  if (owner >= 0) {
    int r = lower_bound_i32(seeds, n_planes, owner);
    pi = r + 1;
    if (r < n_planes && seeds[r] == owner)
      lb = r + 1;
  }
  if (lb == 0) {
    int r = lower_bound_i32(seeds, n_planes, (int32_t)o);
    if (r < n_planes && seeds[r] == (int32_t)o)
      lb = r + 1;
  }
  pidx[o] = pi;
  label[o] = lb;
}

__global__ void export_list_kernel(const int32_t* __restrict__ pool, const int4* __restrict__ pts, int64_t src_off,
                                   int64_t len, int32_t* __restrict__ out)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < len)
    out[t] = pts[pool[src_off + t]].w;
}

__global__ void paint_kernel(const int32_t* __restrict__ label, const uint16_t* __restrict__ rgb, int64_t n,
                             uint16_t* __restrict__ colors)
{
  int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n)
    return;
  int32_t l = label[o];
  uint16_t r = 0, g = 0, b = 0;
  if (l > 0) {
    r = rgb[3 * (l - 1)];
    g = rgb[3 * (l - 1) + 1];
    b = rgb[3 * (l - 1) + 2];
  }
  colors[3 * o] = r;
  colors[3 * o + 1] = g;
  colors[3 * o + 2] = b;
}

// set_plane_color of a filtered / reordered plane vector (my_function.cpp:268-274): the planes are painted in the
// order listed and a later one overwrites, so a point gets the colour of the LAST listed plane whose pointIdx
// holds it.  rank_of[s] = 1 + highest list position among those planes (0 = none), by atomicMax over the entries.
__global__ void paint_rank_kernel(const int32_t* __restrict__ pool, int64_t src_off, int64_t len, int32_t rank,
                                  int32_t* __restrict__ rank_of)
{
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < len)
    atomicMax(rank_of + pool[src_off + t], rank);
}

__global__ void paint_by_rank_kernel(const int32_t* __restrict__ rank_of, const uint32_t* __restrict__ inv,
                                     const uint16_t* __restrict__ rgb, int64_t n, uint16_t* __restrict__ colors)
{
  int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n)
    return;
  const int32_t r = rank_of[inv[o]];
  uint16_t c0 = 0, c1 = 0, c2 = 0;
  if (r > 0) {
    c0 = rgb[3 * (r - 1)];
    c1 = rgb[3 * (r - 1) + 1];
    c2 = rgb[3 * (r - 1) + 2];
  }
  colors[3 * o] = c0;
  colors[3 * o + 1] = c1;
  colors[3 * o + 2] = c2;
}

// class of the plane a point is labelled with (bseg_plane_classes)
__global__ void point_class_kernel(const int32_t* __restrict__ label, const uint8_t* __restrict__ plane_class, int64_t n,
                                   uint8_t* __restrict__ out)
{
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n)
    return;
  const int32_t l = label[o];
  out[o] = l > 0 ? plane_class[l - 1] : (uint8_t)0;
}

}  // namespace

// host copy of the committed planes, sorted by seed (kept between grow and get_planes)
static std::vector<PlaneRec>& host_planes(bseg_ctx* c)
{
  if (!c->h_planes)
    c->h_planes = new std::vector<PlaneRec>();
  return *static_cast<std::vector<PlaneRec>*>(c->h_planes);
}

void grow_host_free(bseg_ctx* c)
{
  if (c->h_planes)
    delete static_cast<std::vector<PlaneRec>*>(c->h_planes);
  c->h_planes = nullptr;
}

int stage_grow_speculative(bseg_ctx* c, const bseg_params* p, GrowArgs& A);  // grow_spec.cu

void launch_grow_seq(bseg_ctx* c, const GrowArgs& A, bool notify, unsigned long long max_tx, unsigned long long max_steps,
                     int allow_growers)
{
  if (notify)
    grow_seq_kernel<true><<<1, 32, 0, c->stream>>>(A, max_tx, max_steps, allow_growers);
  else
    grow_seq_kernel<false><<<1, 32, 0, c->stream>>>(A, max_tx, max_steps, 1);
}

int stage_grow(bseg_ctx* c, const bseg_params* p)
{
  const int64_t n = c->n;
  c->n_planes = 0;
  c->n_plane_entries = 0;
  host_planes(c).clear();
  if (n == 0)
    return 0;
  const int64_t planes_cap = n / (p->th_point_count > 0 ? p->th_point_count : 1) + 16;
  const int64_t pool_cap = 2 * n + planes_cap + 64;
  RC_CHECK(dev_ensure(c, c->g_state, (size_t)n * 8));
  RC_CHECK(dev_ensure(c, c->g_pool, (size_t)pool_cap * 4));
  RC_CHECK(dev_ensure(c, c->g_stack, (size_t)(n + 64) * 8));
  RC_CHECK(dev_ensure(c, c->g_planes, (size_t)planes_cap * sizeof(PlaneRec)));
  RC_CHECK(dev_ensure(c, c->g_label, (size_t)n * 4));
  RC_CHECK(dev_ensure(c, c->g_pidx, (size_t)n * 4));
  RC_CHECK(dev_ensure(c, c->counters, 192 * sizeof(uint64_t)));
  c->pool_cap = pool_cap;

  GrowArgs A;
  A.pts = dptr<int4>(c->pts);
  A.nrm = dptr<double>(c->nrm);
  A.nbr = dptr<int32_t>(c->nbr);
  A.inv = dptr<uint32_t>(c->inv);
  A.state = dptr<int32_t>(c->g_state);                // owner and reservation of point p side by side: state[2p], res[2p]
  A.res = dptr<uint32_t>(c->g_state) + 1;             // (one 8-byte record, one sector for both gathers)
  A.n = n;
  A.K = p->K;
  A.th_thick = (double)p->th_thickness;
  A.th_dot = p->th_dot;
  A.th_count = p->th_point_count;
  A.pool = dptr<int32_t>(c->g_pool);
  A.pool_cap = pool_cap;
  A.stack = dptr<int2>(c->g_stack);
  A.planes = dptr<PlaneRec>(c->g_planes);
  A.planes_cap = planes_cap;
  A.ctl = reinterpret_cast<unsigned long long*>(dptr<uint64_t>(c->counters)) + 64;  // [64, 128): grower control block
  A.doom = nullptr;
  A.slotof = nullptr;
  A.atby = nullptr;
  A.stop_flag = nullptr;
  A.slice_min_ns = 0;
  A.frontier = 0;
  {
    const char* gf = getenv("BSEG_GROW_FLAGS");  // tuning switches of the step engine (grow.cuh GF_*)
    A.flags = gf ? atoi(gf) : (GF_ROWDUP | GF_FASTDIV | GF_ROW_L1);
    if (getenv("BSEG_DEBUG")) A.flags |= GF_TIMING;
    if (c->normals_nonunit) A.flags |= GF_EXACT_MODEL;  // the margins of the approximate model assume unit normals  // per-phase cycle counters of the head slot (they cost ~10 %)
  }
  RC_CHECK(dev_ensure(c, c->g_rowdup, (size_t)n + 64));
  A.rowdup = dptr<uint8_t>(c->g_rowdup);

  STAGE_BEGIN(c, EV_GROW);
  if (p->grow_radius > 0.0) {
    RC_CHECK(dev_ensure(c, c->g_nbr_masked, (size_t)n * p->K * 4));
    const double r2d = p->grow_radius * p->grow_radius;
    unsigned long long r2 = (unsigned long long)r2d;
    if ((double)r2 < r2d) ++r2;  // d2 < r2  <=>  d2 < ceil(r2) for integer d2
    mask_rows_kernel<<<(unsigned)ceil_div64(n * p->K, 256), 256, 0, c->stream>>>(dptr<int32_t>(c->nbr), dptr<int4>(c->pts), n, p->K,
                                                                              r2, dptr<int32_t>(c->g_nbr_masked));
    KLAUNCH_CHECK(c);
    A.nbr = dptr<int32_t>(c->g_nbr_masked);
  }
  if (c->grow_off[0] | c->grow_off[1] | c->grow_off[2]) {
    RC_CHECK(dev_ensure(c, c->g_pts_raw, (size_t)n * 16));
    offset_pts_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(dptr<int4>(c->pts), n, c->grow_off[0], c->grow_off[1],
                                                                          c->grow_off[2], dptr<int4>(c->g_pts_raw));
    KLAUNCH_CHECK(c);
    A.pts = dptr<int4>(c->g_pts_raw);
  }
  rowdup_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(A.nbr, A.K, n, dptr<uint8_t>(c->g_rowdup));
  KLAUNCH_CHECK(c);
  CU_CHECK(c, cudaMemsetAsync(A.state, 0xff, (size_t)n * 8, c->stream));  // state = -1 (free), res = RES_FREE
  CU_CHECK(c, cudaMemsetAsync(A.ctl, 0, 64 * sizeof(unsigned long long), c->stream));
  int64_t rounds = 0;
  if (p->grow_mode == 1) {
    launch_grow_seq(c, A, false, ~0ull, ~0ull, 1);
    KLAUNCH_CHECK(c);
    rounds = 1;
  } else {
    RC_CHECK(stage_grow_speculative(c, p, A));
    rounds = c->tm.grow_rounds;
  }
  STAGE_END(c, EV_GROW);

  unsigned long long ctl[8];
  RC_CHECK(read_back(c, ctl, A.ctl, sizeof(ctl)));
  if (ctl[CTL_ERR])
    return bseg_fail(c, BSEG_E_CAPACITY, "plane grower ran out of list/plane capacity (code %llu)", ctl[CTL_ERR]);
  if ((int64_t)ctl[CTL_FRONTIER] < n)
    return bseg_fail(c, BSEG_E_STATE, "plane grower stopped at seed %llu of %lld", ctl[CTL_FRONTIER], (long long)n);
  c->n_planes = (int32_t)ctl[CTL_PLANES];
  c->tm.grow_steps = (int64_t)ctl[CTL_STEPS];
  c->tm.grow_rounds = rounds;

  // ---- finalize: ordinal plane ids (seed order), planeIdx and labels ----
  STAGE_BEGIN(c, EV_FINALIZE);
  std::vector<PlaneRec>& hp = host_planes(c);
  hp.resize((size_t)c->n_planes);
  if (c->n_planes > 0) {
    CU_CHECK(c, cudaMemcpyAsync(hp.data(), A.planes, hp.size() * sizeof(PlaneRec), cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    std::sort(hp.begin(), hp.end(), [](const PlaneRec& a, const PlaneRec& b) { return a.seed < b.seed; });
  }
  std::vector<int32_t> seeds(hp.size() + 1);
  int64_t entries = 0;
  for (size_t i = 0; i < hp.size(); ++i) {
    seeds[i] = hp[i].seed;
    entries += hp[i].len;
  }
  c->n_plane_entries = entries;
  RC_CHECK(dev_ensure(c, c->g_queue, (seeds.size() + 4) * 4));
  if (!hp.empty())
    CU_CHECK(c, cudaMemcpyAsync(c->g_queue.p, seeds.data(), hp.size() * 4, cudaMemcpyHostToDevice, c->stream));
  finalize_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(A.state, A.inv, dptr<int32_t>(c->g_queue),
                                                                      c->n_planes, n, dptr<int32_t>(c->g_pidx),
                                                                      dptr<int32_t>(c->g_label));
  KLAUNCH_CHECK(c);
  CU_CHECK(c, cudaStreamSynchronize(c->stream));  // `seeds` must outlive the upload
  STAGE_END(c, EV_FINALIZE);
  return 0;
}

int stage_export_grow(bseg_ctx* c, int32_t* h_plane_idx, int32_t* h_label)
{
  const int64_t n = c->n;
  if (n == 0)
    return 0;
  STAGE_BEGIN(c, EV_D2H);
  if (h_plane_idx)
    CU_CHECK(c, cudaMemcpyAsync(h_plane_idx, c->g_pidx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  if (h_label)
    CU_CHECK(c, cudaMemcpyAsync(h_label, c->g_label.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  STAGE_END(c, EV_D2H);
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int stage_get_planes(bseg_ctx* c, int32_t* seeds, double* normals, int32_t* centers, int64_t* offsets,
                     int32_t* point_idx)
{
  std::vector<PlaneRec>& hp = host_planes(c);
  int64_t off = 0;
  for (size_t i = 0; i < hp.size(); ++i) {
    if (seeds) seeds[i] = hp[i].seed;
    if (normals)
      for (int k = 0; k < 3; ++k) normals[3 * i + k] = hp[i].nrm[k];
    if (centers)
      for (int k = 0; k < 3; ++k) centers[3 * i + k] = hp[i].ctr[k];
    if (offsets) offsets[i] = off;
    off += hp[i].len;
  }
  if (offsets) offsets[hp.size()] = off;
  if (point_idx && off > 0) {
    RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)off * 4));
    int64_t o = 0;
    for (size_t i = 0; i < hp.size(); ++i) {
      export_list_kernel<<<(unsigned)ceil_div64(hp[i].len, 256), 256, 0, c->stream>>>(
          dptr<int32_t>(c->g_pool), dptr<int4>(c->pts), hp[i].off, hp[i].len, dptr<int32_t>(c->out_tmp) + o);
      KLAUNCH_CHECK(c);
      o += hp[i].len;
    }
    CU_CHECK(c, cudaMemcpyAsync(point_idx, c->out_tmp.p, (size_t)off * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int stage_paint(bseg_ctx* c, const int32_t* h_ids, int32_t n_listed, const uint16_t* h_rgb, uint16_t* h_colors)
{
  const int64_t n = c->n;
  if (n == 0)
    return 0;
  const size_t rgb_bytes = (size_t)n_listed * 6;
  RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)n * 6 + rgb_bytes + (h_ids ? (size_t)n * 4 : 0) + 128));
  uint16_t* d_colors = dptr<uint16_t>(c->out_tmp);
  uint16_t* d_rgb = d_colors + 3 * n + 8;
  if (rgb_bytes)
    CU_CHECK(c, cudaMemcpyAsync(d_rgb, h_rgb, rgb_bytes, cudaMemcpyHostToDevice, c->stream));
  if (!h_ids) {
    // all planes in id order: the label (last plane holding the point) names the colour
    paint_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(dptr<int32_t>(c->g_label), d_rgb, n, d_colors);
    KLAUNCH_CHECK(c);
  } else {
    std::vector<PlaneRec>& hp = host_planes(c);
    int32_t* d_rank = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(d_rgb) + ((rgb_bytes + 15) & ~(size_t)15));
    CU_CHECK(c, cudaMemsetAsync(d_rank, 0, (size_t)n * 4, c->stream));
    for (int32_t q = 0; q < n_listed; ++q) {
      const PlaneRec& r = hp[(size_t)h_ids[q] - 1];
      paint_rank_kernel<<<(unsigned)ceil_div64(r.len, 256), 256, 0, c->stream>>>(dptr<int32_t>(c->g_pool), r.off, r.len, q + 1,
                                                                                d_rank);
      KLAUNCH_CHECK(c);
    }
    paint_by_rank_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(d_rank, dptr<uint32_t>(c->inv), d_rgb, n, d_colors);
    KLAUNCH_CHECK(c);
  }
  CU_CHECK(c, cudaMemcpyAsync(h_colors, d_colors, (size_t)n * 6, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// DetectedPlane (my_function.h:41-46) of every committed plane + roof / facade / ground classes (include/bseg.h)
int stage_plane_classes(bseg_ctx* c, double facade_max_nz, double roof_min_nz, double ground_z, double* h_eq,
                        uint8_t* h_plane_class, uint8_t* h_point_class)
{
  std::vector<PlaneRec>& hp = host_planes(c);
  std::vector<uint8_t> cls(hp.size() + 1, 0);
  for (size_t i = 0; i < hp.size(); ++i) {
    const double a = hp[i].nrm[0], b = hp[i].nrm[1], cc = hp[i].nrm[2];
    const double cx = (double)hp[i].ctr[0], cy = (double)hp[i].ctr[1], cz = (double)hp[i].ctr[2];
    if (h_eq) {
      h_eq[4 * i] = a;
      h_eq[4 * i + 1] = b;
      h_eq[4 * i + 2] = cc;
      h_eq[4 * i + 3] = -((a * cx + b * cy) + cc * cz);
    }
    const double az = bseg_fabs(cc);
    uint8_t k = BSEG_CLASS_OTHER;
    if (az <= facade_max_nz) k = BSEG_CLASS_FACADE;
    else if (az >= roof_min_nz) k = cz >= ground_z ? BSEG_CLASS_ROOF : BSEG_CLASS_GROUND;
    cls[i] = k;
    if (h_plane_class) h_plane_class[i] = k;
  }
  const int64_t n = c->n;
  if (!h_point_class || n == 0)
    return 0;
  RC_CHECK(dev_ensure(c, c->out_tmp, (size_t)n + cls.size() + 64));
  uint8_t* d_out = dptr<uint8_t>(c->out_tmp);
  uint8_t* d_cls = d_out + ((n + 15) & ~(int64_t)15);
  CU_CHECK(c, cudaMemcpyAsync(d_cls, cls.data(), cls.size(), cudaMemcpyHostToDevice, c->stream));
  point_class_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, c->stream>>>(dptr<int32_t>(c->g_label), d_cls, n, d_out);
  KLAUNCH_CHECK(c);
  CU_CHECK(c, cudaMemcpyAsync(h_point_class, d_out, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}
