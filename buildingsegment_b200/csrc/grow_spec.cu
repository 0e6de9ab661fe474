// grow_spec.cu -- the speculative, order-faithful plane grower (grow_mode 0).
//
// The reference (my_function.cpp:180-258) is a sequential greedy algorithm: seeds in index order,
// each transaction TX(i) = "if point i is still free, run Broad(i,0) and its DFS, then commit or
// roll back".  A transaction's decisions depend on other transactions only through the free/taken
// state of the points it ACCEPTS (a geometric rejection does not look at the state, and a taken
// point stays taken), so TX(i) computed against a snapshot equals its sequential outcome iff no
// lower-index transaction that was still in flight accepted one of the same points or TX(i)'s seed.
//
// Engine (rounds over a window [F, F+C) of original indices, F = first unresolved seed):
//   depth0   one thread per window index: skip taken seeds; run the depth-0 tests; tiny
//            transactions (depth-0 failure: <= K-2 orphan marks) reserve their accepted points with
//            atomicMin(res[pt], i); depth-0 successes become grower candidates
//   assign   the LOWEST candidates get the free grower slots (scan over the window)
//   grow     one warp per slot runs Broad steps for a time slice, reserving every accepted point;
//            slots persist across rounds (a plane may need many slices)
//   overlap detection is symmetric: whoever comes second at a point sees the other's reservation
//            (atomicMin result): the lower index keeps the point, the higher one is doomed
//   validate first_bad = lowest window index whose transaction is doomed, blocked or unfinished
//   commit   everything below first_bad commits in parallel (disjoint point sets by construction);
//            everything at/after it releases its reservations, except slots still running clean
//   head     the sequential engine (grow.cu) then runs transactions from F in order -- this clears
//            the doomed head and gives at least sequential-GPU speed on dependent chains
// Only a PREFIX ever commits, and a waiting transaction is doomed as soon as any lower one touches
// its points, so the final state is exactly the sequential one.  Plane ids are ordinal in seed order
// and are assigned after the fact (grow.cu finalize).
#include "grow.cuh"

namespace {

constexpr int TPB = 256;
constexpr int GW = 4;  // warps per block in slot kernels

enum { KIND_NONE = 0, KIND_SKIP = 1, KIND_TINY = 2, KIND_TINY_BAD = 3, KIND_CAND = 4, KIND_BLOCKED = 5, KIND_SLOT = 8 };
enum { ST_FREE = 0, ST_NEW = 1, ST_RUNNING = 2, ST_FINISHED = 3, ST_DOOMED = 4 };
// spec control block (A.ctl + 8)
enum { SC_FIRST_BAD = 0, SC_NFREE = 1, SC_TX_COMMIT = 2, SC_NCAND = 3, SC_SLOT_STEPS = 4 };

struct Slot {
  int32_t seed_i;
  uint32_t seed_s;
  int32_t status;
  int32_t failed;    // depth-0 failure inside a slot: commits as orphan marks, no plane
  int32_t n_pages;   // pages of the pool this slot owns
  int32_t pad;
  unsigned long long steps;
  TxState t;
};

struct SpecArgs {
  GrowArgs A;
  Slot* slots;
  int G;
  PagePool pool;
  uint32_t* ptabs;  // [G][MAX_PAGES_PER_SLOT]
  uint32_t n_pool_pages;
  uint32_t* kind;   // [C] window-relative
  uint32_t* mask;   // [C]
  uint32_t* flag;   // [C] candidate flags -> exclusive scan
  uint32_t* free_ids;
  unsigned long long* sc;
};

__device__ __forceinline__ PagedStore slot_store(const SpecArgs& S, int g)
{
  PagedStore st;
  st.pool = S.pool;
  st.ptab = S.ptabs + (size_t)g * MAX_PAGES_PER_SLOT;
  st.n_pages = &S.slots[g].n_pages;
  return st;
}

__global__ void __launch_bounds__(TPB) spec_prepare_kernel(SpecArgs S, int64_t F, int64_t C)
{
  int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (t >= C)
    return;
  S.kind[t] = KIND_NONE;
  S.flag[t] = 0;
  if (!S.A.hasslot[F + t])
    S.A.doom[F + t] = 0;
}

__global__ void __launch_bounds__(TPB) spec_mark_slots_kernel(SpecArgs S, int64_t F, int64_t C)
{
  int g = blockIdx.x * TPB + threadIdx.x;
  if (g >= S.G)
    return;
  const Slot& sl = S.slots[g];
  if (sl.status == ST_FREE)
    return;
  int64_t t = (int64_t)sl.seed_i - F;
  if (t >= 0 && t < C)
    S.kind[t] = KIND_SLOT + (uint32_t)g;
}

// one thread per window index: Broad(i, 0) against the committed state (my_function.cpp:221-239)
__global__ void __launch_bounds__(TPB) spec_depth0_kernel(SpecArgs S, int64_t F, int64_t C)
{
  const GrowArgs& A = S.A;
  int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (t >= C || S.kind[t] != KIND_NONE)
    return;
  const int64_t i = F + t;
  const uint32_t s = __ldg(A.inv + i);
  if (__ldcg(A.state + s) != -1) {
    S.kind[t] = KIND_SKIP;
    return;
  }
  if (__ldcg(A.res + s) < (uint32_t)i) {  // a lower in-flight transaction accepted this seed
    S.kind[t] = KIND_BLOCKED;
    return;
  }
  Model m;
  model_init(m, __ldg(A.pts + s), A.nrm + 3 * (int64_t)s);
  const int K = A.K;
  const int32_t* row = A.nbr + (int64_t)s * K;
  uint32_t mask = 0;
  int cnt = 0;
  for (int j = 1; j < K; ++j) {
    const int32_t id = __ldg(row + j);
    if (id < 0 || __ldcg(A.state + id) != -1)
      continue;
    bool dup = false;
    for (int j2 = 1; j2 < j; ++j2)
      if (((mask >> j2) & 1u) && __ldg(row + j2) == id) dup = true;
    if (dup)
      continue;
    const int4 p = __ldg(A.pts + id);
    const double* nr = A.nrm + 3 * (int64_t)id;
    if (geo_test(m, p, nr[0], nr[1], nr[2], A.th_thick, A.th_dot)) {
      mask |= 1u << j;
      ++cnt;
    }
  }
  if (cnt == K - 1) {
    S.kind[t] = KIND_CAND;
    S.flag[t] = 1;
    return;
  }
  bool lost = false;
  for (int j = 1; j < K; ++j)
    if ((mask >> j) & 1u) {
      const int32_t id = __ldg(row + j);
      const uint32_t old = atomicMin(A.res + id, (uint32_t)i);
      if (old < (uint32_t)i) lost = true;
      else {
        if (old != RES_FREE && old != (uint32_t)i) A.doom[old] = 1;
        const int32_t o = __ldg(A.pts + id).w;
        if (o > (int32_t)i) A.doom[o] = 1;
      }
    }
  S.mask[t] = mask;
  S.kind[t] = lost ? KIND_TINY_BAD : KIND_TINY;
}

// flag[] holds the exclusive scan of the candidate flags; the lowest candidates take the free slots
__global__ void __launch_bounds__(TPB) spec_assign_kernel(SpecArgs S, int64_t F, int64_t C)
{
  int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (t >= C || S.kind[t] != KIND_CAND)
    return;
  const uint32_t r = S.flag[t];
  const uint32_t nfree = (uint32_t)S.sc[SC_NFREE];
  if (r >= nfree) {
    S.kind[t] = KIND_BLOCKED;
    return;
  }
  const uint32_t g = S.free_ids[nfree - 1 - r];
  Slot& sl = S.slots[g];
  sl.seed_i = (int32_t)(F + t);
  sl.seed_s = __ldg(S.A.inv + F + t);
  sl.status = ST_NEW;
  sl.steps = 0;
  sl.t.len = 0;
  sl.failed = 0;
  sl.n_pages = 0;
  S.A.hasslot[F + t] = 1;
  S.kind[t] = KIND_SLOT + g;
}

__global__ void spec_pop_free_kernel(SpecArgs S)
{
  const unsigned long long nc = S.sc[SC_NCAND] & 0xffffffffull, nf = S.sc[SC_NFREE];
  S.sc[SC_NFREE] = nf - (nc < nf ? nc : nf);
}

// one warp per slot: a time slice of Broad steps with reservations
__global__ void __launch_bounds__(GW * 32) spec_grow_kernel(SpecArgs S, unsigned long long budget)
{
  const GrowArgs& A = S.A;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * GW + (threadIdx.x >> 5);
  if (g >= S.G)
    return;
  Slot& sl = S.slots[g];
  const int status = sl.status;
  if (status != ST_NEW && status != ST_RUNNING)
    return;
  const int64_t seed_i = sl.seed_i;
  if (A.doom[seed_i]) {
    __syncwarp();
    if (lane == 0) sl.status = ST_DOOMED;  // len stays 0 for a slot that never started
    return;
  }
  PagedStore st = slot_store(S, g);
  TxState t;
  if (status == ST_NEW) {
    tx_begin(t, A, sl.seed_s);
    if (!st.reserve(1, lane)) {
      if (lane == 0) sl.status = ST_DOOMED;
      return;
    }
    if (lane == 0) st.put(0, (int32_t)sl.seed_s);
    __syncwarp();
  } else {
    t = sl.t;
  }
  __syncwarp();
  unsigned long long steps = 0;
  const TxOutcome out = tx_run<MODE_SPEC>(A, st, t, seed_i, budget, false, lane, steps);
  __syncwarp();
  if (lane == 0) {
    sl.t = t;
    sl.steps += steps;
    sl.failed = out == TX_FAILED ? 1 : 0;
    sl.status = out == TX_RUNNING ? ST_RUNNING : ((out == TX_FINISHED || out == TX_FAILED) ? ST_FINISHED : ST_DOOMED);
  }
}

__global__ void __launch_bounds__(TPB) spec_validate_kernel(SpecArgs S, int64_t F, int64_t C)
{
  int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  bool bad = false;
  if (t < C) {
    const uint32_t k = S.kind[t];
    if (k == KIND_SKIP) bad = false;
    else if (k == KIND_TINY) bad = S.A.doom[F + t] != 0;
    else if (k >= KIND_SLOT) bad = S.slots[k - KIND_SLOT].status != ST_FINISHED || S.A.doom[F + t] != 0;
    else bad = true;
  }
  // lowest bad index of the block -> one atomic
  __shared__ unsigned long long blk;
  if (threadIdx.x == 0) blk = ~0ull;
  __syncthreads();
  if (bad) atomicMin(&blk, (unsigned long long)t);
  __syncthreads();
  if (threadIdx.x == 0 && blk != ~0ull) atomicMin(&S.sc[SC_FIRST_BAD], blk);
}

__global__ void __launch_bounds__(TPB) spec_commit_tiny_kernel(SpecArgs S, int64_t F, int64_t C)
{
  const GrowArgs& A = S.A;
  int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (t >= C)
    return;
  const uint32_t k = S.kind[t];
  if (k != KIND_TINY && k != KIND_TINY_BAD)
    return;
  const unsigned long long fbu = S.sc[SC_FIRST_BAD];
  const int64_t first_bad = fbu > (unsigned long long)C ? C : (int64_t)fbu;
  const int64_t i = F + t;
  const uint32_t s = __ldg(A.inv + i);
  const int32_t* row = A.nbr + (int64_t)s * A.K;
  const uint32_t mask = S.mask[t];
  if (t < first_bad) {
    for (int j = 1; j < A.K; ++j)
      if ((mask >> j) & 1u) A.state[__ldg(row + j)] = (int32_t)i;  // orphan marks, :233 then :238-239
    atomicAdd(&S.sc[SC_TX_COMMIT], 1ull);
  } else {
    for (int j = 1; j < A.K; ++j)
      if ((mask >> j) & 1u) atomicCAS(A.res + __ldg(row + j), (uint32_t)i, RES_FREE);
  }
}

__global__ void __launch_bounds__(GW * 32) spec_commit_slots_kernel(SpecArgs S, int64_t F, int64_t C)
{
  const GrowArgs& A = S.A;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * GW + (threadIdx.x >> 5);
  if (g >= S.G)
    return;
  Slot& sl = S.slots[g];
  if (sl.status == ST_FREE)
    return;
  const unsigned long long fbu = S.sc[SC_FIRST_BAD];
  const int64_t first_bad = fbu > (unsigned long long)C ? C : (int64_t)fbu;  // slots beyond the window were not validated
  const int64_t i = sl.seed_i;
  const int64_t t = i - F;
  const PagedStore st = slot_store(S, g);
  const int64_t len = sl.t.len;
  bool release = false, free_slot = false;
  if (t < first_bad) {  // finished and clean
    free_slot = true;
    if (sl.failed) {
      for (int64_t e = 1 + lane; e < len; e += 32) A.state[st.get(e)] = (int32_t)i;
    } else if (len > A.th_count) {
      unsigned long long off = 0, pl = 0;
      if (lane == 0) {
        off = atomicAdd(&A.ctl[CTL_POOL], (unsigned long long)len);
        pl = atomicAdd(&A.ctl[CTL_PLANES], 1ull);
      }
      off = __shfl_sync(FULL_MASK, off, 0);
      pl = __shfl_sync(FULL_MASK, pl, 0);
      if ((int64_t)(off + len) > A.pool_cap - A.n - 2 || (int64_t)pl >= A.planes_cap) {
        if (lane == 0) A.ctl[CTL_ERR] = 2;
      } else {
        for (int64_t e = lane; e < len; e += 32) {
          const int32_t id = st.get(e);
          A.pool[off + e] = id;
          if (e >= 1) A.state[id] = (int32_t)i;
        }
        if (lane == 0) {
          PlaneRec r;
          r.seed = (int32_t)i; r.pad = 0;
          r.off = (int64_t)off; r.len = len;
          r.nrm[0] = sl.t.m.mn0; r.nrm[1] = sl.t.m.mn1; r.nrm[2] = sl.t.m.mn2;
          r.ctr[0] = sl.t.m.mc0; r.ctr[1] = sl.t.m.mc1; r.ctr[2] = sl.t.m.mc2; r.pad2 = 0;
          A.planes[pl] = r;
        }
      }
    } else {
      release = true;  // :203-209 roll back: nothing persists
    }
    if (lane == 0) {
      atomicAdd(&S.sc[SC_TX_COMMIT], 1ull);
      atomicAdd(&A.ctl[CTL_STEPS], sl.steps);
    }
  } else if (sl.status == ST_DOOMED || A.doom[i]) {
    release = true;
    free_slot = true;
    if (lane == 0) atomicAdd(&S.sc[SC_SLOT_STEPS], sl.steps);
  }
  if (release)
    for (int64_t e = lane; e < len; e += 32) atomicCAS(A.res + st.get(e), (uint32_t)i, RES_FREE);
  __syncwarp();
  if (free_slot && lane == 0) {
    // pages back to the pool
    const int np = sl.n_pages;
    if (np > 0) {
      const unsigned long long pos = atomicAdd(S.pool.n_free, (unsigned long long)np);
      for (int k = 0; k < np; ++k) S.pool.free_pages[pos + k] = st.ptab[k];
    }
    sl.n_pages = 0;
    sl.status = ST_FREE;
    A.hasslot[i] = 0;
    const unsigned long long pos = atomicAdd(&S.sc[SC_NFREE], 1ull);
    S.free_ids[pos] = (uint32_t)g;
  }
}

__global__ void spec_advance_kernel(SpecArgs S, int64_t F, int64_t C)
{
  unsigned long long fb = S.sc[SC_FIRST_BAD];
  if (fb > (unsigned long long)C) fb = (unsigned long long)C;
  const unsigned long long Fn = (unsigned long long)F + fb;
  S.A.ctl[CTL_FRONTIER] = Fn;
  S.sc[5] = fb;  // reported to the host
  S.sc[6] = (Fn < (unsigned long long)S.A.n && S.A.hasslot[Fn]) ? 1ull : 0ull;  // a live slot sits at the head
}

__global__ void spec_init_kernel(SpecArgs S)
{
  uint32_t g = blockIdx.x * TPB + threadIdx.x;
  if (g < (uint32_t)S.G) {
    S.slots[g].status = ST_FREE;
    S.slots[g].n_pages = 0;
    S.free_ids[g] = (uint32_t)(S.G - 1 - g);
  }
  if (g < S.n_pool_pages) S.pool.free_pages[g] = g;
  if (g == 0) {
    S.sc[SC_NFREE] = (unsigned long long)S.G;
    *S.pool.n_free = (unsigned long long)S.n_pool_pages;
  }
}

}  // namespace

void launch_grow_seq(bseg_ctx* c, const GrowArgs& A, bool notify, unsigned long long max_tx, unsigned long long max_steps,
                     int allow_growers);

int stage_grow_speculative(bseg_ctx* c, const bseg_params* p, GrowArgs& A)
{
  (void)p;
  const int64_t n = A.n;
  const int64_t CMAX = 1 << 20, CMIN = 4096;
  SpecArgs S;
  S.G = 1024;
  // page pool: room for every point once plus one page per slot, capped at 64 Mi entries
  int64_t pages = (3 * n) / PAGE_SIZE + 2 * S.G;
  if (pages > 32768) pages = 32768;
  S.n_pool_pages = (uint32_t)pages;
  const size_t slot_bytes = (size_t)S.G * sizeof(Slot) + (size_t)S.G * 4 + 256;
  RC_CHECK(dev_ensure(c, c->g_tx, slot_bytes + (size_t)S.G * MAX_PAGES_PER_SLOT * 4 + (size_t)pages * 4 + 64));
  RC_CHECK(dev_ensure(c, c->g_spec, (size_t)CMAX * 12 + (size_t)n * 2 + 256));
  RC_CHECK(dev_ensure(c, c->g_queue, (size_t)pages * PAGE_SIZE * 12 + 256));
  S.slots = dptr<Slot>(c->g_tx);
  S.free_ids = reinterpret_cast<uint32_t*>(S.slots + S.G);
  S.ptabs = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(c->g_tx.p) + slot_bytes);
  S.pool.free_pages = S.ptabs + (size_t)S.G * MAX_PAGES_PER_SLOT;
  S.pool.stack_pages = dptr<int2>(c->g_queue);
  S.pool.list_pages = reinterpret_cast<int32_t*>(S.pool.stack_pages + (size_t)pages * PAGE_SIZE);
  S.kind = dptr<uint32_t>(c->g_spec);
  S.mask = S.kind + CMAX;
  S.flag = S.mask + CMAX;
  A.doom = reinterpret_cast<uint8_t*>(S.flag + CMAX);
  A.hasslot = A.doom + n;
  S.sc = A.ctl + 8;
  S.pool.n_free = &S.sc[7];
  S.A = A;
  CU_CHECK(c, cudaMemsetAsync(A.doom, 0, (size_t)n * 2, c->stream));
  {
    const int64_t m = pages > S.G ? pages : S.G;
    spec_init_kernel<<<(unsigned)ceil_div64(m, TPB), TPB, 0, c->stream>>>(S);
    KLAUNCH_CHECK(c);
  }

  int64_t F = 0, C = 16384, rounds = 0;
  unsigned long long head_tx = 64;
  int64_t seq_growers = 0;
  uint32_t* d_ncand = reinterpret_cast<uint32_t*>(&S.sc[SC_NCAND]);
  while (F < n) {
    if (C > n - F) C = n - F;
    const unsigned gb = (unsigned)ceil_div64(C, TPB);
    const unsigned sb = (unsigned)((S.G + GW - 1) / GW);
    CU_CHECK(c, cudaMemsetAsync(&S.sc[SC_FIRST_BAD], 0xff, sizeof(unsigned long long), c->stream));
    spec_prepare_kernel<<<gb, TPB, 0, c->stream>>>(S, F, C);
    KLAUNCH_CHECK(c);
    spec_mark_slots_kernel<<<(S.G + TPB - 1) / TPB, TPB, 0, c->stream>>>(S, F, C);
    KLAUNCH_CHECK(c);
    spec_depth0_kernel<<<gb, TPB, 0, c->stream>>>(S, F, C);
    KLAUNCH_CHECK(c);
    CU_CHECK(c, cudaMemsetAsync(&S.sc[SC_NCAND], 0, sizeof(unsigned long long), c->stream));
    RC_CHECK(bseg_exclusive_scan_u32(c, S.flag, C, d_ncand));
    spec_assign_kernel<<<gb, TPB, 0, c->stream>>>(S, F, C);
    KLAUNCH_CHECK(c);
    spec_pop_free_kernel<<<1, 1, 0, c->stream>>>(S);
    KLAUNCH_CHECK(c);
    spec_grow_kernel<<<sb, GW * 32, 0, c->stream>>>(S, 4096ull);
    KLAUNCH_CHECK(c);
    spec_validate_kernel<<<gb, TPB, 0, c->stream>>>(S, F, C);
    KLAUNCH_CHECK(c);
    spec_commit_tiny_kernel<<<gb, TPB, 0, c->stream>>>(S, F, C);
    KLAUNCH_CHECK(c);
    spec_commit_slots_kernel<<<sb, GW * 32, 0, c->stream>>>(S, F, C);
    KLAUNCH_CHECK(c);
    spec_advance_kernel<<<1, 1, 0, c->stream>>>(S, F, C);
    KLAUNCH_CHECK(c);
    launch_grow_seq(c, S.A, true, head_tx, 1ull << 40, 0);
    KLAUNCH_CHECK(c);
    ++rounds;
    unsigned long long ctl[16];
    RC_CHECK(read_back(c, ctl, A.ctl, sizeof(ctl)));
    if (ctl[CTL_ERR])
      break;
    const int64_t committed = (int64_t)ctl[8 + 5];
    const bool head_live = ctl[8 + 6] != 0;
    int64_t Fn = (int64_t)ctl[CTL_FRONTIER];
    if (Fn < F)
      return bseg_fail(c, BSEG_E_STATE, "speculative grower: frontier moved backwards");
    if (Fn == F && !head_live) {
      // the head is a depth-0 success without a slot (no slot / no pages left, or a plane too large for
      // the page table): the head runner grows it in the flat region, alone
      launch_grow_seq(c, S.A, true, 1, 1ull << 40, 1);
      KLAUNCH_CHECK(c);
      RC_CHECK(read_back(c, ctl, A.ctl, sizeof(ctl)));
      if (ctl[CTL_ERR])
        break;
      Fn = (int64_t)ctl[CTL_FRONTIER];
      ++seq_growers;
    }
    F = Fn;
    // window: wide when the clean prefix is long, never below CMIN (speculation is cheap)
    if (committed >= C) C = C * 2 < CMAX ? C * 2 : CMAX;
    else {
      int64_t want = committed * 4;
      C = want < CMIN ? CMIN : (want > CMAX ? CMAX : want);
    }
    // head runner: long sequential runs where the window keeps colliding, a single step otherwise
    head_tx = committed < 256 ? 2048 : (committed < 4096 ? 128 : 8);
    if (rounds > (int64_t)4 * n + 1024)
      return bseg_fail(c, BSEG_E_STATE, "speculative grower: no progress");
  }
  c->tm.grow_rounds = rounds;
  (void)seq_growers;
  return 0;
}
