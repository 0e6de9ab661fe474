// grow_spec.cu -- the parallel, order-faithful plane grower (grow_mode 0): a sequential SWEEPER that is
// the only writer of the committed state, fed by growers that run AHEAD of it speculatively.
//
// The reference (my_function.cpp:180-258) is a sequential greedy algorithm: seeds in index order, each
// transaction TX(i) = "if point i is still free, run Broad(i,0) and its DFS, then commit or roll back".
// Two facts make it parallel without changing a single label:
//   (1) Whether a neighbour passes the depth-0 geometric tests depends only on the seed's own normal and
//       position (my_function.cpp:189-190,227-230), never on the state.  gmask[s] (one bit per neighbour
//       column, duplicates removed) is therefore computed for ALL points up front, and a seed whose
//       transaction fails at depth 0 (:238-239, the overwhelming majority: "tiny" transactions that only
//       leave orphan marks) costs the sweeper one state gather.  The sweeper is ONE thread block that walks
//       the seeds in index order, 1024 at a time: every thread evaluates its seed against the committed
//       state, a shared-memory hash table finds the seeds that a lower seed of the same batch is about to
//       mark, and the conflict-free prefix of the batch commits in parallel (atomicMin keeps the lower owner
//       where two tiny transactions want the same point).
//   (2) A transaction's decisions depend on other transactions only through the free/taken state of the
//       points it ACCEPTS (a geometric rejection does not look at the state, a taken point stays taken).
//       Depth-0 successes ("growers") therefore run ahead of the sweeper in slots, one warp each, against
//       the committed state, reserving every accepted point with atomicMin(res[pt], seed index); tiny
//       transactions of the window ahead of the sweeper publish the marks they are going to make the same
//       way (hints, re-evaluated and retracted every round).  A point that is
//         - reserved by a LOWER in-flight transaction is treated as taken and appended to the grower's
//           assumed-taken (AT) list;
//         - reserved by a HIGHER one is taken, and a higher grower that held it is doomed;
//         - marked by the sweeper while a grower holds it: the grower is doomed.
//       When the sweeper reaches a seed that is a grower at its turn, every lower transaction has been
//       committed; a finished, un-doomed slot whose AT points are all taken by then holds exactly the
//       sequential result and is committed (or rolled back, :203-209) in place.  Otherwise the sweeper
//       stops, the seed gets a slot as the lowest candidate (one slot is kept for it) and -- with nothing
//       lower in flight -- runs clean: progress is guaranteed.
// Round = release doomed slots -> scout (lowest candidates of the window take the free slots) -> one slice
// of Broad steps for every running slot (ends when the head slot finishes) -> verification of the finished
// slots -> sweep, with a BACKGROUND slice of the growers and an L2 warmer beside it on the other SMs -> the
// side effects of the sweep (marks, accepted planes) applied by the whole GPU.  Every kernel of a round reads
// the frontier from the control block, so the host launches round r + 1 before it has seen round r's result
// (stage_grow_speculative).  Plane ids are ordinal in seed order and are assigned after the fact (grow.cu
// finalize).
#include <cstdlib>

#include "grow.cuh"

namespace {

constexpr int TPB = 256;
constexpr int GW = 4;          // warps per block in slot kernels
constexpr int SWEEP_T = 512;   // seeds per sweeper batch == threads of the sweeper block (128 registers each: ids and reservations stay in registers)
constexpr int SWEEP_WORDS = 128;  // bitmap words (4096 seeds) scanned per batch for up to SWEEP_T live seeds

constexpr int PEND_CAP = 64;       // assumed-taken points of a finished slot that were still free when it was verified
constexpr int CSET_BITS = 11, CSET = 1 << CSET_BITS;  // sweeper: hash set of the seeds whose planes this sweep committed (<= G of them)
constexpr int GSET_BITS = 11, GSET = 1 << GSET_BITS;  // sweeper: seeds of the grower-looking threads of a segment (<= SWEEP_T of them)
constexpr int DSET_BITS = 10, DSET = 1 << DSET_BITS;  // sweeper: seeds it doomed during this sweep

// DONE: slot and pages still to be returned; COMMIT: the sweeper accepted the plane, its list is still in the slot
// (spec_apply_commits_kernel copies it out); ROLLED: rolled back at its turn, the reservations are still to be dropped
enum { ST_FREE = 0, ST_RUNNING = 1, ST_FINISHED = 2, ST_DEAD = 3, ST_RELEASING = 4, ST_DONE = 5, ST_COMMIT = 6, ST_ROLLED = 7 };
// spec control block (A.ctl + 8)
enum {
  SC_NFREE = 0,       // free slots
  SC_NCAND = 1,       // candidates of the window (scan total)
  SC_STOP = 44,       // head finished: slices end (1 = now, 2 = once the slice is old enough); [45] = start of the slice (globaltimer)
  SC_HEAD_SLOT = 3,   // slot of the seed at the frontier + 1 (0 = none)
  SC_WASTED = 4,      // Broad steps of released slots
  SC_SWEEP_ITERS = 5,
  SC_TINY = 6,        // tiny transactions committed
  SC_POOLFREE = 7,    // free pages
  SC_NREL = 8,        // released seeds of this round
  SC_STUCK = 9,       // sweeper stopped at a grower that has no slot
  SC_NASSIGN = 10,    // slots handed out by this round's scout
  SC_HEAD_STEPS = 11, // Broad steps of the head slot (the critical path) ...
  SC_HEAD_NS = 12,    // ... and the time they took (globaltimer)
  SC_SWEEP_NS = 13,   // time inside the sweeper
  SC_ATFAIL = 14,     // finished slots whose assumed-taken points were not all taken
  SC_NLOG = 19,       // entries of the mark log
  SC_SWEEP_PH = 46,   // [46, 53): sweeper sub-step phases (ns, thread 0)
  SC_SWEEP_ON = 41,   // the sweeper of this round is running (background slice: the growers of the other SMs work while it does)
  SC_BG_STEPS = 42,   // Broad steps made by background slices
  SC_LIVE_F = 43,     // the sweeper's frontier, published once per batch (the prefetch block walks ahead of it)
  SC_END_NOSLOT = 53, SC_END_RUNNING = 54, SC_END_UNVERIFIED = 55,  // why sweeps ended: the grower at the frontier has no slot / is still growing / finished beside the sweeper
  SC_N_SER = 22,      // serial stretches of the sweeper
  SC_N_GROW = 23,     // growers the sweeper decided
  SC_HEAD_ITERS = 21, // warp iterations of the head slot (two-node engine: <= steps)
  SC_POOL0 = 20,      // pool fill at the start of the sweep (planes committed since: [SC_POOL0, CTL_POOL))
  SC_SKIP_DBG = 24,   // [24, 29): head skip batches, cycles in enumerate / row+state / geometry, pairs tested
  SC_T_FRONT = 15, SC_T_SLOW = 16, SC_T_FAST = 17, SC_N_SLOW = 18,  // sweeper time split (ns), slow-path count
};

__device__ __forceinline__ unsigned long long gtimer()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct Slot {
  int32_t seed_i;
  uint32_t seed_s;
  int32_t status;
  int32_t started;   // 0 = tx_begin still to be done
  int32_t n_pages;   // pages of the pool this slot owns
  int32_t n_pend;    // verified slots: assumed-taken points that were still free then (first PEND_CAP in S.pend)
  int32_t verified;  // spec_preverify_kernel has checked the list of this finished slot
  int32_t pad;
  int64_t pool_off;  // ST_COMMIT: where the list goes in the committed pool
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
  const uint32_t* gmask;  // [n] ORIGINAL index space (the sweeper and the scout walk seeds in index order)
  uint32_t* flag;   // [CW] candidate bit << 31 | rank of the candidate inside its scout block
  uint32_t* bsum;   // [blocks of the scout] candidates per block
  uint32_t* free_ids;
  uint8_t* hinted;    // [n] original index space: this tiny transaction has published hints
  uint32_t* alive;    // [ceil(n/32)] original index space: bit = the point may still be free (filter)
  int2* pend;         // [G][PEND_CAP] see Slot::n_pend: (point, who held it when the slot was verified)
  uint2* marklog;     // (point, seed) of the orphan marks the sweeper made this round: side effects applied later
  unsigned long long marklog_cap;
  unsigned long long* sc;
};

__device__ __forceinline__ PagedStore slot_store(const SpecArgs& S, int g)
{
  PagedStore st;
  st.pool = S.pool;
  st.ptab = S.ptabs + (size_t)g * MAX_PAGES_PER_SLOT;
  st.n_pages = &S.slots[g].n_pages;
  st.have_cached = S.slots[g].n_pages;
  return st;
}

// ---- depth-0 geometry of every point (state-independent) -------------------------------------------------
// bit j of gmask[s]: neighbour column j of s is a valid point, passes the tests of Broad(s, 0) with the
// seed's own model (my_function.cpp:189-190,227-230) and is the first such occurrence of that point in the
// row (a repeated id finds planeIdx already set, :226).
__global__ void __launch_bounds__(TPB) gmask_kernel(GrowArgs A, uint32_t* __restrict__ gmask)
{
  const int64_t s = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (s >= A.n)
    return;
  Model m;
  const int4 self = __ldg(A.pts + s);
  model_init(m, self, A.nrm + 3 * s);
  const int K = A.K;
  const int32_t* row = A.nbr + s * K;
  uint32_t mask = 0;
  for (int j = 1; j < K; ++j) {
    const int32_t id = __ldg(row + j);
    if (id < 0)
      continue;
    bool dup = false;
    for (int j2 = 1; j2 < j; ++j2)
      if (((mask >> j2) & 1u) && __ldg(row + j2) == id) dup = true;
    if (dup)
      continue;
    const int4 p = __ldg(A.pts + id);
    const double* nr = A.nrm + 3 * (int64_t)id;
    if (geo_test(m, p, __ldg(nr), __ldg(nr + 1), __ldg(nr + 2), A.th_thick, A.th_dot)) mask |= 1u << j;
  }
  gmask[self.w] = mask;
}

// ---- slot bookkeeping shared by the release kernel and the sweeper ----------------------------------------
// give the slot's pages and the slot itself back (one thread)
__device__ __forceinline__ void slot_free(const SpecArgs& S, int g)
{
  Slot& sl = S.slots[g];
  const PagedStore st = slot_store(S, g);
  const int np = sl.n_pages;
  if (np > 0) {
    const unsigned long long pos = atomicAdd(S.pool.n_free, (unsigned long long)np);
    for (int k = 0; k < np; ++k) S.pool.free_pages[pos + k] = st.ptab[k];
  }
  sl.n_pages = 0;
  sl.status = ST_FREE;
  S.A.slotof[sl.seed_i] = -1;
  const unsigned long long pos = atomicAdd(&S.sc[SC_NFREE], 1ull);
  S.free_ids[pos] = (uint32_t)g;
}

// ---- K0: release doomed / dead slots ahead of the sweeper, then doom whoever relied on them -----------------
// A doomed plane can hold 10^5 reservations: RCH blocks per slot walk its list.
constexpr int RCH = 4;  // (x G slots: the grid is launched every round and nearly all of its blocks have nothing to do)
// which slots go: decided ONCE (releasing a slot can doom others through the early notification, and the
// three kernels must agree)
__global__ void __launch_bounds__(TPB) spec_mark_release_kernel(SpecArgs S)
{
  const int g = blockIdx.x * TPB + threadIdx.x;
  if (g >= S.G)
    return;
  Slot& sl = S.slots[g];
  if (sl.status == ST_FREE || sl.status == ST_DONE || sl.status == ST_ROLLED)
    return;
  if (sl.status == ST_COMMIT) {  // its list has been copied out by spec_apply_commits_kernel
    sl.status = ST_DONE;
    return;
  }
  if (sl.status == ST_DEAD || S.A.doom[sl.seed_i]) sl.status = ST_RELEASING;
  else if (sl.status == ST_FINISHED) {
    sl.n_pend = 0;  // recounted by this round's spec_preverify_kernel
    if (sl.verified == 1) sl.verified = 2;  // its list was walked last round: only the open assumptions from now on
  }
}

__global__ void __launch_bounds__(TPB) spec_release_entries_kernel(SpecArgs S)
{
  const GrowArgs& A = S.A;
  const int g = blockIdx.y;
  const Slot& sl = S.slots[g];
  if (sl.status != ST_RELEASING && sl.status != ST_ROLLED)  // ROLLED: my_function.cpp:203-209 at its turn, nothing persists
    return;
  const int32_t i = sl.seed_i;
  const PagedStore st = slot_store(S, g);
  const int64_t len = sl.started ? sl.t.len : 0;
  for (int64_t e = 1 + (int64_t)blockIdx.x * TPB + threadIdx.x; e < len; e += (int64_t)RCH * TPB) {
    const int32_t pt = st.get(e);
    if (atomicCAS(A.res + 2 * (int64_t)(pt), (uint32_t)i, RES_FREE) == (uint32_t)i) unreserve_notify(A.atby, A.slotof, A.doom, pt);
  }
}

__global__ void __launch_bounds__(TPB) spec_release_slots_kernel(SpecArgs S)
{
  const int g = blockIdx.x * TPB + threadIdx.x;
  if (g >= S.G)
    return;
  Slot& sl = S.slots[g];
  if (sl.status != ST_RELEASING && sl.status != ST_DONE && sl.status != ST_ROLLED)
    return;
  if (sl.status == ST_RELEASING) atomicAdd(&S.sc[SC_WASTED], sl.steps);
  slot_free(S, g);
}

// ---- side effects of the sweeper's orphan marks (see spec_sweep_kernel) ------------------------------------------
__global__ void __launch_bounds__(TPB) spec_apply_marks_kernel(SpecArgs S)
{
  const GrowArgs& A = S.A;
  unsigned long long nlog = S.sc[SC_NLOG];
  if (nlog > S.marklog_cap) nlog = S.marklog_cap;
  for (unsigned long long k = (unsigned long long)blockIdx.x * TPB + threadIdx.x; k < nlog;
       k += (unsigned long long)gridDim.x * TPB) {
    const uint2 e = S.marklog[k];
    if (e.x == 0xffffffffu)
      continue;  // log space a serial stretch reserved and did not need
    const uint32_t r = __ldcg(A.res + 2 * (int64_t)(e.x));
    if (r != RES_FREE) {
      if (r != e.y) A.doom[r] = 1;  // a grower ahead of the sweeper held it: void
      A.res[2 * (int64_t)e.x] = RES_FREE;        // (a hint of this or a higher tiny transaction: obsolete)
    }
    const int32_t w = __ldg(&A.pts[e.x].w);
    atomicAnd(S.alive + (w >> 5), ~(1u << (w & 31)));
    if (__ldcg(A.slotof + w) >= 0) A.doom[w] = 1;  // the grower seeded at this point lost its seed
  }
}

// ---- finished slots: the list is checked ONCE, in parallel, when the slot has finished ----------------------------
// A finished, un-doomed slot holds the sequential result iff at its turn (V1) every point it accepted is still free
// and (V2) every point it assumed taken is taken.  V1 holds at its turn iff it holds now AND the slot still owns every
// reservation AND nothing marks one of its points later -- and whoever does that (a lower grower stealing the
// reservation, the scout publishing a lower hint, the sweeper marking the point for a tiny transaction) dooms the
// holder at that moment.  So the list is walked here by the whole GPU instead of by the sweeper's one block; V2 is
// reduced to the (few) assumed-taken points that are still free now, which the sweeper looks at when the turn comes.
__global__ void __launch_bounds__(TPB) spec_preverify_kernel(SpecArgs S)
{
  const GrowArgs& A = S.A;
  const int g = blockIdx.y;
  if (blockIdx.x == 0 && g == 0 && threadIdx.x == 0) {  // the background slice that runs beside the sweeper: ends when the sweeper does
    S.sc[SC_STOP] = 0;
    S.sc[SC_SWEEP_ON] = 0;
    S.sc[SC_LIVE_F] = 0;
  }
  Slot& sl = S.slots[g];
  if (sl.status != ST_FINISHED)
    return;
  const int32_t i = sl.seed_i;
  if (((volatile uint8_t*)A.doom)[i])
    return;
  const PagedStore st = slot_store(S, g);
  // the list once (verified: 0 -> 1 here, 1 -> 2 by the next round's spec_mark_release_kernel), the open assumptions every round
  const int64_t len = sl.verified == 2 ? 0 : sl.t.len, n_at = sl.t.n_at;
  bool bad = false;
  for (int64_t e = 1 + (int64_t)blockIdx.x * TPB + threadIdx.x; e < len; e += (int64_t)RCH * TPB) {
    const int32_t pt = st.get(e);
    const int2 sr = ld_state_res(A, pt);
    bad |= sr.x != -1 || (uint32_t)sr.y != (uint32_t)i;
  }
  for (int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x; k < n_at; k += (int64_t)RCH * TPB) {
    const int32_t pt = st.get_at(k);
    const int2 sr = ld_state_res(A, pt);
    if (sr.x == -1) {
      const int idx = atomicAdd(&sl.n_pend, 1);
      if (idx < PEND_CAP) S.pend[(size_t)g * PEND_CAP + idx] = make_int2(pt, sr.y);
    }
  }
  if (bad) A.doom[i] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0 && sl.verified == 0) sl.verified = 1;  // (not 2: the other blocks of the slot test it)
}

// ---- planes the sweep accepted: list -> committed pool, owner marks, reservations dropped, alive bits ---------------
// (the sweeper itself only records the plane and names its seed in the set of committed seeds: for the rest of the
// sweep a point reserved by such a seed counts as taken)
__global__ void __launch_bounds__(TPB) spec_apply_commits_kernel(SpecArgs S)
{
  const GrowArgs& A = S.A;
  const int g = blockIdx.y;
  if (blockIdx.x == 0 && g == 0 && threadIdx.x == 0) {  // the mark log has been applied (spec_apply_marks_kernel): empty it
    S.sc[SC_NLOG] = 0;
    S.sc[SC_POOL0] = A.ctl[CTL_POOL];
  }
  const Slot& sl = S.slots[g];
  if (sl.status != ST_COMMIT)
    return;
  const int32_t i = sl.seed_i;
  const PagedStore st = slot_store(S, g);
  const int64_t len = sl.t.len, off = sl.pool_off;
  for (int64_t e = (int64_t)blockIdx.x * TPB + threadIdx.x; e < len; e += (int64_t)RCH * TPB) {
    const int32_t id = st.get(e);
    A.pool[off + e] = id;
    if (e >= 1) {  // the seed's own entry is not a mark (:191)
      A.state[2 * (int64_t)id] = i;
      A.res[2 * (int64_t)id] = RES_FREE;
    }
    const int32_t w = __ldg(&A.pts[id].w);
    atomicAnd(S.alive + (w >> 5), ~(1u << (w & 31)));
  }
}

// ---- K1: scout -- the window [F, F+C) ahead of the sweeper -------------------------------------------------
// Every seed of the window without a slot is evaluated against the committed state plus the reservations
// of LOWER in-flight transactions (reservations below F are stale and count as free):
//   dead (seed taken, or reserved by a lower transaction): retract the hints it may have published;
//   grower-looking (all K-1 neighbours pass depth 0, free and unreserved by anything lower): candidate;
//   tiny-looking: publish the marks it is going to make, atomicMin(res[pt], i) -- a higher grower that
//     held the point is doomed.
__device__ __forceinline__ uint32_t eff_res(uint32_t r, uint32_t F) { return r < F ? RES_FREE : r; }

// The frontier F is read from the control block (the host runs a round ahead of what it has seen); the window is
// [F, F + min(n - F, CW)), the grid always covers CW seeds.
__global__ void __launch_bounds__(TPB) spec_scout_kernel(SpecArgs S, int64_t CW)
{
  const GrowArgs& A = S.A;
  const int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  const int64_t F = (int64_t)__ldcg(A.ctl + CTL_FRONTIER);
  const int64_t C = A.n - F < CW ? A.n - F : CW;
  if (t == 0) {
    S.sc[SC_STOP] = 0;
    S.sc[SC_NCAND] = 0;
    S.sc[SC_NASSIGN] = 0;
  }
  const int64_t i = F + t;
  const uint32_t me = (uint32_t)i, fr = (uint32_t)F;
  uint32_t cand = 0;
  if (t < C) {
  // the alive bitmap is a filter (bit clear => the point is taken): late in the pass five seeds of six end here, for one
  // coalesced word and one byte instead of three loads and a gather
  const bool hinted = S.hinted[i] != 0;
  const bool maybe_free = ((__ldcg(S.alive + (i >> 5)) >> (int)(i & 31)) & 1u) != 0;
  if ((maybe_free || hinted) && __ldcg(A.slotof + i) < 0) {
    const uint32_t s = __ldg(A.inv + i);
    const int K = A.K;
    const uint32_t m = __ldg(S.gmask + i);
    const int32_t* row = A.nbr + (int64_t)s * K;
    const int2 sr_s = ld_state_res(A, s);
    const bool dead = sr_s.x != -1 || eff_res((uint32_t)sr_s.y, fr) < me;
    uint32_t want = 0;
    bool lower = false;
    if (!dead) {
      uint32_t mm = m;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const int32_t id = __ldg(row + j);
        const int2 sr = ld_state_res(A, id);
        if (sr.x == -1) {
          want |= 1u << j;
          lower |= eff_res((uint32_t)sr.y, fr) < me;
        }
      }
      cand = (__popc(m) == K - 1 && __popc(want) == K - 1 && !lower) ? 1u : 0u;
    }
    if (dead || cand) {
      if (hinted) {  // retract: a grower must not find its own stale hints, a dead seed marks nothing
        uint32_t mm = m;
        while (mm) {
          const int j = __ffs(mm) - 1;
          mm &= mm - 1;
          const int32_t pt = __ldg(row + j);
          if (atomicCAS(A.res + 2 * (int64_t)(pt), me, RES_FREE) == me) unreserve_notify(A.atby, A.slotof, A.doom, pt);
        }
        S.hinted[i] = 0;
      }
    } else if (want) {
      uint32_t mm = want;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const int32_t id = __ldg(row + j);
        uint32_t cur = __ldcg(A.res + 2 * (int64_t)(id));
        for (;;) {
          if (cur != RES_FREE && cur >= fr && cur <= me)
            break;  // ours already, or a lower transaction's
          uint32_t prev;
          if (cur < fr) {  // stale: replace
            prev = atomicCAS(A.res + 2 * (int64_t)(id), cur, me);
            if (prev != cur) {
              cur = prev;
              continue;
            }
          } else {
            prev = atomicMin(A.res + 2 * (int64_t)(id), me);
            if (prev < fr) {  // cannot happen inside this kernel (nobody writes stale values); be safe
              cur = prev;
              continue;
            }
            if (prev != RES_FREE && prev > me && __ldcg(A.slotof + prev) >= 0) A.doom[prev] = 1;
          }
          break;
        }
      }
      S.hinted[i] = 1;
    }
  }
  }
  // rank of a candidate among the candidates of its block, and the block's count: spec_assign_kernel adds the counts of
  // the blocks below -- the window's exclusive scan without scan kernels (three launches per round)
  __shared__ uint32_t wsum[TPB / 32];
  const uint32_t bal = __ballot_sync(FULL_MASK, cand != 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wsum[warp] = (uint32_t)__popc(bal);
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < TPB / 32; ++w) {
    const uint32_t cw = wsum[w];
    if (w < warp) before += cw;
    total += cw;
  }
  if (t < CW) S.flag[t] = (before + (uint32_t)__popc(bal & lanemask_lt())) | (cand << 31);
  if (threadIdx.x == 0) S.bsum[blockIdx.x] = total;
}

// flag[t] = candidate bit << 31 | rank inside the scout's block, bsum[b] = candidates of block b: the lowest candidates
// of the window take the free slots
__global__ void __launch_bounds__(TPB) spec_assign_kernel(SpecArgs S, int64_t CW)
{
  const int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  const int64_t F = (int64_t)__ldcg(S.A.ctl + CTL_FRONTIER);
  const int64_t C = S.A.n - F < CW ? S.A.n - F : CW;
  // candidates in the blocks below this one
  __shared__ uint32_t wpart[TPB / 32];
  uint32_t part = 0;
  for (uint32_t k = threadIdx.x; k < blockIdx.x; k += TPB) part += S.bsum[k];
  for (int o = 16; o; o >>= 1) part += __shfl_down_sync(FULL_MASK, part, o);
  if ((threadIdx.x & 31) == 0) wpart[threadIdx.x >> 5] = part;
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int w = 0; w < TPB / 32; ++w) base += wpart[w];
  if (t >= C)
    return;
  const uint32_t f = S.flag[t];
  if (!(f >> 31))
    return;  // not a candidate
  const uint32_t r = base + (f & 0x7fffffffu);
  // the last free slot is kept for the seed at the frontier: the head can always grow
  const uint32_t nfree = (uint32_t)S.sc[SC_NFREE];
  if (t == 0 ? r >= nfree : r + 1 >= nfree)
    return;
  atomicAdd(&S.sc[SC_NASSIGN], 1ull);
  const uint32_t g = S.free_ids[nfree - 1 - r];
  Slot& sl = S.slots[g];
  sl.seed_i = (int32_t)(F + t);
  sl.seed_s = __ldg(S.A.inv + F + t);
  sl.status = ST_RUNNING;
  sl.started = 0;
  sl.steps = 0;
  sl.n_pages = 0;
  sl.n_pend = 0;
  sl.verified = 0;
  S.A.slotof[F + t] = (int32_t)g;
  S.A.doom[F + t] = 0;
}

__global__ void spec_pop_free_kernel(SpecArgs S)
{
  const int64_t F = (int64_t)S.A.ctl[CTL_FRONTIER];
  S.sc[SC_NFREE] -= S.sc[SC_NASSIGN];
  S.sc[SC_STOP + 1] = gtimer();  // the slice starts
  const int32_t g = F < S.A.n ? S.A.slotof[F] : -1;
  S.sc[SC_HEAD_SLOT] = (unsigned long long)(g + 1);
  // the head finished in a background slice: nothing to wait for, verification and the sweeper are next
  if (g >= 0 && S.slots[g].status != ST_RUNNING) S.sc[SC_STOP] = 1;
}

// ---- K2: one warp per slot, a time slice of Broad steps with reservations ----------------------------------
// bg = 1: the BACKGROUND slice.  The sweeper is one block on one SM; while it walks the seeds the slots keep growing on
// the other SMs.  Nothing the sweeper decides depends on a slot that is still running (it stops at the first grower
// that is not finished AND verified), a running slot reads the committed state (which only gets more taken: reading an
// old value is reading earlier) and reserves FREE points with atomics, and whatever the sweeper marks under a running
// slot's reservation is found by spec_apply_marks_kernel after both have ended (the holder is doomed) and again by the
// full-list verification of the slot.  What a running slot must NOT do beside the sweeper is take a reservation away
// from its holder: the sweeper decides a finished slot by its flags (read at the front of the batch), and recognises
// the points of a plane it has accepted in this sweep by the reservations the plane still holds -- a slot the
// sweeper has just voided can run on for a few steps, and a reservation it stole in that time would make a taken point
// look free.  With GF_BG a call that would have to steal ends the slot's slice instead (grow.cuh, tx_run_pair).  The slice ends when the sweeper sets SC_STOP; a slot that does not see the
// sweeper running within a few microseconds (it could not be placed, or has already ended) does nothing.
__global__ void __launch_bounds__(GW * 32) spec_grow_kernel(SpecArgs S, unsigned long long budget, int bg)
{
  const GrowArgs& A = S.A;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * GW + (threadIdx.x >> 5);
  if (g >= S.G)
    return;
  Slot& sl = S.slots[g];
  if (sl.status != ST_RUNNING)
    return;
  if (bg) {
    const unsigned long long t_wait = gtimer();
    for (;;) {
      if (*(volatile unsigned long long*)&S.sc[SC_STOP] != 0)
        return;
      if (*(volatile unsigned long long*)&S.sc[SC_SWEEP_ON] != 0)
        break;
      if (gtimer() - t_wait > 20000ull)
        return;
      __nanosleep(200);
    }
  }
  const int64_t seed_i = sl.seed_i;
  if (((volatile uint8_t*)A.doom)[seed_i]) {
    __syncwarp();
    if (lane == 0) sl.status = ST_DEAD;
    return;
  }
  PagedStore st = slot_store(S, g);
  TxState t;
  if (!sl.started) {
    tx_begin(t, A, sl.seed_s);
    if (!st.reserve(1, lane)) {
      if (lane == 0) sl.status = ST_DEAD;
      return;
    }
    if (lane == 0) st.put(0, (int32_t)sl.seed_s);
    __syncwarp();
  } else {
    t = sl.t;
  }
  __syncwarp();
  unsigned long long steps = 0, iters = 0;
  const bool is_head = !bg && (unsigned long long)(g + 1) == S.sc[SC_HEAD_SLOT];
  const unsigned long long t0 = is_head ? gtimer() : 0ull;
  // K <= 16: two DFS nodes per warp step (grow.cuh tx_run_pair); BSEG_GROW_FLAGS bit 6 forces the single-node engine
  const bool pair = A.K <= 16 && !(A.flags & GF_NOPAIR);
  __shared__ SkipScratch skip_scratch[GW];
  SkipScratch* ss = &skip_scratch[threadIdx.x >> 5];
  unsigned long long* dbg = (is_head && (A.flags & GF_TIMING)) ? &S.sc[SC_SKIP_DBG] : nullptr;
  const TxOutcome out = pair ? (A.K == 15 ? tx_run_pair<15>(A, st, t, seed_i, budget, lane, steps, &iters, ss, dbg)
                                          : tx_run_pair<0>(A, st, t, seed_i, budget, lane, steps, &iters, ss, dbg))
                             : tx_run<MODE_SPEC, 0>(A, st, t, seed_i, budget, false, lane, steps);
  __syncwarp();
  if (lane == 0 && is_head) {
    S.sc[SC_HEAD_STEPS] += steps;
    S.sc[SC_HEAD_ITERS] += iters;
    S.sc[SC_HEAD_NS] += gtimer() - t0;
  }
  if (lane == 0 && bg) atomicAdd(&S.sc[SC_BG_STEPS], steps);
  if (lane == 0) {
    sl.t = t;
    sl.started = 1;
    sl.steps += steps;
    // a depth-0 failure under assumptions is a tiny transaction: the sweeper's job
    sl.status = out == TX_RUNNING ? ST_RUNNING : (out == TX_FINISHED ? ST_FINISHED : ST_DEAD);
    // the head sets the pace: when it finishes OR has used its budget the slice is over for everybody
    // (a slot that keeps running after the head has stopped only delays the sweeper)
    if (is_head) {
      // ... unless it was a short one: then the others run on until the slice is slice_min_ns old
      const bool early = out != TX_RUNNING && gtimer() - S.sc[SC_STOP + 1] < A.slice_min_ns;
      *(volatile unsigned long long*)&S.sc[SC_STOP] = early ? 2ull : 1ull;
    }
  }
}

// ---- K3: the sweeper ------------------------------------------------------------------------------------------
// One thread block walks the seeds in index order, SWEEP_T live seeds per batch.  A batch is resolved in sub-steps; one
// sub-step decides a SEGMENT [done, seg_end) of the batch -- tiny transactions and growers together -- in parallel:
//   * every live seed is classified against the committed state: tiny-looking (marks its free depth-0 neighbours) or
//     grower-looking (all K-1 are free: :238 does not fire);
//   * a shared-memory hash table maps every point a tiny-looking seed wants to the lowest thread wanting it, a second
//     one maps the seeds of the grower-looking threads to their threads;
//   * a seed is in CONFLICT when its outcome depends on a lower seed of the same segment in a way the tables cannot
//     settle: its own point is wanted by a lower tiny seed or reserved by a lower grower-looking one; a grower-looking
//     seed has a neighbour wanted by a lower tiny seed; any of its wanted points is reserved by a lower grower-looking
//     seed.  The segment ends at the first conflict (the seed is re-classified after the segment has committed);
//   * tiny seeds of the segment doom the holder of every point they are about to mark (a grower ahead of the sweeper),
//     then the grower-looking seeds decide, each on its own thread: the slot must be finished, verified
//     (spec_preverify_kernel) and un-doomed, and the assumed-taken points that were still free at verification must
//     be taken now.  The first grower that cannot be decided ends the segment AND the sweep;
//   * commit: orphan marks (atomicMin on the owner: the lower seed keeps a shared point) and the mark log; a plane is
//     only RECORDED (pool range, PlaneRec, status) and its seed enters the set of committed seeds -- for the rest of
//     the sweep a point reserved by a committed seed counts as taken; the list is copied out by the whole GPU after
//     the sweep (spec_apply_commits_kernel).  A roll-back (:203-209) is a status change.
// Clouds in scan order make nearly every seed depend on the one before it; there the block switches to SERIAL
// STRETCHES: the seeds are resolved one after the other, a warp at a time, against a shared-memory set of the points
// marked so far in the stretch (~200 cycles per live seed instead of one block-wide sub-step per conflict).
// Nothing on the block's path waits for global memory except one gather level per sub-step (states + reservations).
struct SweepShared {
  int first_over;     // lowest thread beyond the hash-table budget
  int first_conf;     // lowest thread in conflict with a lower seed of the segment
  int first_fail;     // lowest grower-looking thread that cannot be decided (the sweep stops there)
  int first_g;        // lowest unfinished grower-looking thread
  int ser_end, ser_ins;  // serial stretch: first thread it did not reach, points in the taken set
  int n_cset, n_dset, dset_over;
  int n_items;
  unsigned long long tph[12], tlast;  // BSEG_DEBUG: thread 0's clock per phase of the sub-step
  unsigned long long log_base;
  int n_conf;         // threads of the segment in conflict
  int sp_big, sp_bad, sp_slot;
  unsigned long long c_off, c_pl, c_steps, c_tx;
  int warp_sum[32];
  int n_live;
  int64_t last_seed, last_open;
  uint32_t words[SWEEP_WORDS];
  int pref[SWEEP_WORDS];
  uint32_t cset[CSET];    // seeds whose planes this sweep committed (open addressing, 0xffffffff = empty)
  uint32_t dset[DSET];    // seeds this sweep doomed
  uint32_t gkey[GSET];    // seeds of the grower-looking threads of the segment ...
  uint32_t gval[GSET];    // ... and their threads
};

template <int BITS>
__device__ __forceinline__ bool set_has(const uint32_t* tab, uint32_t key)
{
  uint32_t h = (key * 2654435761u) >> (32 - BITS);
  for (;;) {
    const uint32_t k = tab[h];
    if (k == key)
      return true;
    if (k == 0xffffffffu)
      return false;
    h = (h + 1) & ((1u << BITS) - 1);
  }
}

// true = the key was not there
template <int BITS>
__device__ __forceinline__ bool set_insert(uint32_t* tab, uint32_t key)
{
  uint32_t h = (key * 2654435761u) >> (32 - BITS);
  for (;;) {
    const uint32_t old = atomicCAS(tab + h, 0xffffffffu, key);
    if (old == 0xffffffffu)
      return true;
    if (old == key)
      return false;
    h = (h + 1) & ((1u << BITS) - 1);
  }
}

// thread of the grower-looking seed `seed` in this segment, SWEEP_T if there is none
__device__ __forceinline__ uint32_t gset_find(const uint32_t* gkey, const uint32_t* gval, uint32_t seed)
{
  uint32_t h = (seed * 2654435761u) >> (32 - GSET_BITS);
  for (;;) {
    const uint32_t k = gkey[h];
    if (k == seed)
      return gval[h];
    if (k == 0xffffffffu)
      return (uint32_t)SWEEP_T;
    h = (h + 1) & (GSET - 1);
  }
}

template <int KM>
struct SweepCfg {
  static constexpr int ITEMS = 4096;  // wants of the tiny seeds of one segment (flat list: 12 B each)
  static constexpr int HT_BITS = 13;  // hash table slots of a segment (keys + vals = 64 KB of shared memory): load <= 0.5
  static constexpr int HT = 1 << HT_BITS;
  static constexpr size_t SMEM = (size_t)HT * 8 + (size_t)ITEMS * 12 + (size_t)SWEEP_T * 4;
};

// the sequential authority: walks the seeds from the frontier in index order
template <int KM>
__global__ void __launch_bounds__(SWEEP_T) spec_sweep_kernel(SpecArgs S)
{
  constexpr int HT_BITS = SweepCfg<KM>::HT_BITS;
  constexpr int HT = SweepCfg<KM>::HT;
  extern __shared__ uint32_t smem_u32[];
  uint32_t* hkeys = smem_u32;        // [HT] point, 0xffffffff = empty
  uint32_t* hvals = smem_u32 + HT;   // [HT] lowest thread of the segment that wants it
  constexpr int ITEMS = SweepCfg<KM>::ITEMS;
  uint32_t* it_id = smem_u32 + 2 * HT;             // [ITEMS] flat list of the segment's wants: point ...
  uint32_t* it_r = it_id + ITEMS;                  // ... its reservation as gathered ...
  uint32_t* sh_me = it_r + ITEMS;                  // [SWEEP_T] seed of every thread
  uint16_t* it_tid = reinterpret_cast<uint16_t*>(sh_me + SWEEP_T);  // ... the thread that wants it ...
  uint16_t* it_slot = it_tid + ITEMS;              // ... and where it sits in the hash table
  __shared__ SweepShared sh;
  __shared__ int32_t tile[KM][32];  // serial stretches: id columns of the warp whose turn it is
  const GrowArgs& A = S.A;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int K = A.K;
  int64_t F = (int64_t)A.ctl[CTL_FRONTIER];
  const int64_t n_words = (A.n + 31) >> 5;
  uint32_t iters = 0, ntiny = 0;
  const bool timing = (A.flags & GF_TIMING) != 0;  // BSEG_DEBUG: time split of the block (thread 0's clock)
  const unsigned long long t_begin = gtimer();
  if (tid == 0) *(volatile unsigned long long*)&S.sc[SC_SWEEP_ON] = 1ull;  // the background slice may start
  for (int k = tid; k < HT; k += SWEEP_T) {
    hkeys[k] = 0xffffffffu;
    hvals[k] = 0xffffffffu;
  }
  for (int k = tid; k < CSET; k += SWEEP_T) sh.cset[k] = 0xffffffffu;
  for (int k = tid; k < DSET; k += SWEEP_T) sh.dset[k] = 0xffffffffu;
  for (int k = tid; k < GSET; k += SWEEP_T) {
    sh.gkey[k] = 0xffffffffu;
    sh.gval[k] = (uint32_t)SWEEP_T;
  }
  if (tid == 0) {
    sh.n_cset = 0;
    sh.n_dset = 0;
    sh.dset_over = 0;
    sh.c_off = A.ctl[CTL_POOL];
    sh.c_pl = A.ctl[CTL_PLANES];
    sh.c_steps = 0;
    sh.c_tx = 0;
    for (int k = 0; k < 12; ++k) sh.tph[k] = 0;
  }
  __syncthreads();
#define PHASE(k) if (timing && tid == 0) { const unsigned long long now_ = gtimer(); sh.tph[k] += now_ - sh.tlast; sh.tlast = now_; }
  unsigned long long t_front = 0;  // (thread 0 only)
  uint32_t n_sub = 0, n_grow = 0, n_ser = 0, n_pendsum = 0;
  bool stop = false;
  bool serial = false;     // conflicts are dense: resolve stretches serially (re-probed every 8th batch)
  bool force_par = false;  // the stretch stopped at a grower: one parallel sub-step decides it

  // this sweep dooms transaction r (a grower ahead of the sweeper that holds a point which is being marked)
  auto doom_now = [&](uint32_t r) {
    A.doom[r] = 1;
    if (sh.dset_over)
      return;
    uint32_t h = (r * 2654435761u) >> (32 - DSET_BITS);
    for (int probe = 0; probe < 32; ++probe) {  // bounded: many threads insert at once
      const uint32_t old = atomicCAS(sh.dset + h, 0xffffffffu, r);
      if (old == r)
        return;
      if (old == 0xffffffffu) {
        if (atomicAdd(&sh.n_dset, 1) >= DSET / 2) sh.dset_over = 1;
        return;
      }
      h = (h + 1) & (DSET - 1);
    }
    sh.dset_over = 1;  // too many to keep: growers then ask global memory
  };

  while (F < A.n && !stop) {
    ++iters;
    if (tid == 0) *(volatile unsigned long long*)&S.sc[SC_LIVE_F] = (unsigned long long)F;
    if ((iters & 7) == 0) serial = false;
    unsigned long long ti0 = 0;
    if (timing && tid == 0) ti0 = gtimer();
    // ---- the next live seeds: scan the alive bitmap (original index order, a superset of the free points) ----
    // Late in the sweep 5 seeds of 6 are taken; they cost one bit here instead of a gather.
    const int64_t base = F & ~31LL;
    if (tid < SWEEP_WORDS) {
      const int64_t wi = (base >> 5) + tid;
      uint32_t w = wi < n_words ? __ldcg(S.alive + wi) : 0u;
      if (tid == 0) w &= 0xffffffffu << (int)(F & 31);  // seeds below the frontier are done
      sh.words[tid] = w;
      sh.pref[tid] = __popc(w);
    }
    __syncthreads();
    if (tid < 32) {  // exclusive scan of SWEEP_WORDS counts by one warp
      int carry = 0;
      for (int k0 = 0; k0 < SWEEP_WORDS; k0 += 32) {
        const int c = sh.pref[k0 + tid];
        int v = c;
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(FULL_MASK, v, o);
          if (tid >= o) v += u;
        }
        sh.pref[k0 + tid] = carry + v - c;
        carry += __shfl_sync(FULL_MASK, v, 31);
      }
      if (tid == 0) sh.n_live = carry;
    }
    __syncthreads();
    const int n_live = sh.n_live;
    const int n_eval = n_live < SWEEP_T ? n_live : SWEEP_T;
    const int64_t stretch_end = (base + 32LL * SWEEP_WORDS < A.n) ? base + 32LL * SWEEP_WORDS : A.n;
    if (n_live == 0) {  // nothing free in this stretch
      F = stretch_end;
      __syncthreads();
      continue;
    }
    // thread k evaluates the k-th live seed of the stretch
    const bool valid = tid < n_eval;
    int64_t i = A.n;
    if (valid) {
      int lo = 0, hi = SWEEP_WORDS - 1;  // last word whose exclusive prefix is <= tid
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sh.pref[mid] <= tid) lo = mid;
        else hi = mid - 1;
      }
      const int bit = __fns(sh.words[lo], 0, tid - sh.pref[lo] + 1);
      i = base + 32LL * lo + bit;
    }
    if (tid == n_eval - 1) sh.last_seed = i;  // the seed after it starts the next batch when all of these are done
    const uint32_t me = (uint32_t)i;
    uint32_t s = 0, want = 0;
    uint32_t held_lo = 0, held_hi = 0;  // wanted neighbour j carries the reservation of a lower / higher seed
    uint32_t rs_self = RES_FREE;
    uint32_t rs[KM];                    // reservation of neighbour j (as gathered by classify)
    int32_t ids[KM];                    // neighbour columns that pass depth 0, -1 = none (registers: every loop over them is unrolled)
#pragma unroll
    for (int j = 1; j < KM; ++j) ids[j] = -1;
    int32_t slot = -1;
    bool live = false, glook = false;
    bool g_ready = false, g_doom0 = false;  // the slot of this seed: finished + verified; doomed before the sweep
    int64_t g_len = 0;
    int g_npend = 0;
    bool pend_done = false, pend_bad = false;  // the block walked this grower's assumed-taken list (more than PEND_CAP were open)
    // committed state + planes this sweep has accepted: one load level (state and reservation of the seed and of its
    // depth-0 neighbours)
    auto classify = [&]() {
      const int2 sr_s = ld_state_res(A, s);
      const int32_t st_s = sr_s.x;  // the bitmap is only a filter: this is the truth
      rs_self = (uint32_t)sr_s.y;
      int32_t stj[KM];
#pragma unroll
      for (int j = 1; j < KM; ++j) {
        stj[j] = 0;
        rs[j] = RES_FREE;
        const int32_t id = ids[j];
        if (id >= 0) {
          const int2 sr = ld_state_res(A, id);
          stj[j] = sr.x;
          rs[j] = (uint32_t)sr.y;
        }
      }
      live = st_s == -1;
      want = 0;
      held_lo = 0;
      held_hi = 0;
      if (live && rs_self != RES_FREE && set_has<CSET_BITS>(sh.cset, rs_self)) live = false;
      if (live) {
#pragma unroll
        for (int j = 1; j < KM; ++j)
          if (stj[j] == -1) {  // (-1 only for a column that exists)
            const uint32_t r = rs[j];
            if (r != RES_FREE && r != me) {
              if (set_has<CSET_BITS>(sh.cset, r))
                continue;  // in a plane this sweep committed: taken
              if (r < me) held_lo |= 1u << j;
              else held_hi |= 1u << j;
            }
            want |= 1u << j;
          }
      }
      glook = live && __popc(want) == K - 1;  // :238 -- every neighbour accepted
    };
    if (valid) {
      // dependent levels of loads: (inv, gmask, slot) -> (row, slot record) -> states
      s = __ldg(A.inv + i);
      const uint32_t m = __ldg(S.gmask + i);
      slot = __ldcg(A.slotof + i);
      const int32_t* row = A.nbr + (int64_t)s * K;
#pragma unroll
      for (int j = 1; j < KM; ++j)
        if ((m >> j) & 1u) ids[j] = __ldg(row + j);
      if (slot >= 0) {  // (its own fields: nobody writes them during the sweep)
        const Slot& sl = S.slots[slot];
        g_ready = sl.status == ST_FINISHED && sl.verified >= 1;
        g_len = sl.t.len;
        g_npend = sl.n_pend;
        g_doom0 = ((volatile uint8_t*)A.doom)[i] != 0;
        if (g_ready) {  // what the decision and the commit read later: towards L1 now
          prefetch_l1(&sl.t.m);
          prefetch_l1(reinterpret_cast<const char*>(&sl.t.m) + 64);
          if (g_npend > 0) {
            const char* pp = reinterpret_cast<const char*>(&S.pend[(size_t)slot * PEND_CAP]);
            prefetch_l1(pp);
            if (g_npend > 16) prefetch_l1(pp + 128);
          }
        }
      }
      classify();
    }
    if (timing && tid == 0) t_front += gtimer() - ti0;

    // ---- sub-steps on this batch: threads [0, done) are finished ----
    int done = 0;
    while (done < n_eval && !stop) {
      ++n_sub;
      if (tid == 0) {
        sh.first_over = SWEEP_T;
        sh.first_conf = SWEEP_T;
        sh.first_fail = SWEEP_T;
        sh.first_g = SWEEP_T;
        sh.n_conf = 0;
        sh.sp_big = 0;
        sh.sp_bad = 0;
        sh.ser_end = n_eval;
        sh.ser_ins = 0;
      }
      __syncthreads();
      const bool act = valid && tid >= done;

      if (serial && !force_par) {
        // ---- serial stretch ----
        // log space for the whole stretch, one entry per want of a live seed (unused ones are voided): ONE reservation
        // in global memory instead of one round trip per warp turn
        int s_woff;  // this thread's first entry, relative to its warp's
        {
          const int nw = (act && live) ? __popc(want) : 0;
          int wincl = nw;
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, wincl, o);
            if (lane >= o) wincl += v;
          }
          s_woff = wincl - nw;
          if (lane == 31) sh.warp_sum[tid >> 5] = wincl;
          __syncthreads();
          if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < SWEEP_T / 32; ++w) {
              const int c = sh.warp_sum[w];
              sh.warp_sum[w] = tot;  // exclusive
              tot += c;
            }
            sh.log_base = tot > 0 ? atomicAdd(&S.sc[SC_NLOG], (unsigned long long)tot) : 0ull;
          }
          __syncthreads();
        }
        for (int wi = done >> 5; wi <= ((n_eval - 1) >> 5); ++wi) {
          if ((tid >> 5) == wi && sh.ser_end == n_eval) {
            const bool logged = act && live;  // (has log space, whatever happens to it below)
            bool mine = logged;
            if (mine && set_has<HT_BITS>(hkeys, s)) {  // marked by an earlier warp of the stretch (:185)
              mine = false;
              if (slot >= 0) A.doom[i] = 1;
            }
            if (mine) {
#pragma unroll
              for (int j = 1; j < KM; ++j) tile[j][lane] = ids[j];
            }
            __syncwarp();
            uint32_t todo = __ballot_sync(FULL_MASK, mine);
            int n_ins = sh.ser_ins;
            const unsigned long long lbase = sh.log_base + (unsigned long long)sh.warp_sum[wi];
            const int woff = s_woff;
            if (logged && !mine) {  // dead before its turn: its entries are void
              int k = 0;
              for (uint32_t b = want; b; b &= b - 1, ++k) {
                const unsigned long long pos = lbase + (unsigned long long)(woff + k);
                if (pos < S.marklog_cap) S.marklog[pos] = make_uint2(0xffffffffu, me);
              }
            }
            int end_tid = -1;
            while (todo) {
              const int l = __ffs(todo) - 1;
              const uint32_t s_l = __shfl_sync(FULL_MASK, s, l);
              const uint32_t want_l = __shfl_sync(FULL_MASK, want, l), hi_l = __shfl_sync(FULL_MASK, held_hi, l);
              const uint32_t seed_l = __shfl_sync(FULL_MASK, me, l);
              const int woff_l = __shfl_sync(FULL_MASK, woff, l);
              const int slot_l = __shfl_sync(FULL_MASK, slot, l);
              const bool glook_l = __shfl_sync(FULL_MASK, (int)glook, l) != 0;
              const bool w = lane >= 1 && lane < KM && ((want_l >> lane) & 1u);
              const int32_t myid = w ? tile[lane][l] : -1;  // lane j looks after neighbour column j of the seed
              const bool dead = set_has<HT_BITS>(hkeys, s_l);  // marked by a lower seed of the stretch (:185)
              if (!dead && glook_l) {
                // still a grower at its turn?  (a neighbour may have been marked in the stretch)
                const bool gone = w && set_has<HT_BITS>(hkeys, (uint32_t)myid);
                if (__popc(__ballot_sync(FULL_MASK, w && !gone)) == K - 1) {
                  end_tid = wi * 32 + l;  // yes: the stretch ends in front of it
                  break;
                }
              }
              todo &= todo - 1;
              if (slot_l >= 0 && lane == 0) A.doom[seed_l] = 1;  // dead, or not a grower at its turn: the slot is void
              bool ins = false;
              if (w && !dead) ins = set_insert<HT_BITS>(hkeys, (uint32_t)myid);
              if (w) {
                const unsigned long long pos = lbase + (unsigned long long)(woff_l + __popc(want_l & lanemask_lt()));
                if (pos < S.marklog_cap) S.marklog[pos] = make_uint2(ins ? (uint32_t)myid : 0xffffffffu, seed_l);
              }
              if (ins) {
                A.state[2 * (int64_t)myid] = (int32_t)seed_l;  // the first marker of the stretch owns the point
                if ((hi_l >> lane) & 1u) {        // held by a higher transaction (rare): it is void
                  const uint32_t r = __ldcg(A.res + 2 * (int64_t)(myid));
                  if (r != RES_FREE && r > seed_l) doom_now(r);
                }
              }
              if (!dead && lane == l) ++ntiny;
              n_ins += __popc(__ballot_sync(FULL_MASK, ins));
              if (n_ins > HT / 2 - 32) {  // the set is full: the stretch ends after this seed
                end_tid = wi * 32 + l + 1;
                break;
              }
            }
            if (end_tid >= 0 && mine && tid >= end_tid) {  // log space of the seeds the warp did not reach
              int k = 0;
              for (uint32_t b = want; b; b &= b - 1, ++k) {
                const unsigned long long pos = lbase + (unsigned long long)(woff + k);
                if (pos < S.marklog_cap) S.marklog[pos] = make_uint2(0xffffffffu, me);
              }
            }
            if (lane == 0) {
              sh.ser_ins = n_ins;
              if (end_tid >= 0 && end_tid < n_eval) sh.ser_end = end_tid;
            }
          }
          __syncthreads();
        }
        const int ser_end = sh.ser_end;
        if (act && live && tid >= ser_end) {  // log space of the seeds the stretch did not reach: void
          const unsigned long long lb = sh.log_base + (unsigned long long)(sh.warp_sum[tid >> 5] + s_woff);
          int k = 0;
          for (uint32_t b = want; b; b &= b - 1, ++k)
            if (lb + k < S.marklog_cap) S.marklog[lb + k] = make_uint2(0xffffffffu, me);
        }
        if (tid == ser_end && valid) sh.last_open = i;
        for (int k = tid; k < HT; k += SWEEP_T) hkeys[k] = 0xffffffffu;
        __syncthreads();
        if (ser_end == done) force_par = true;  // a grower at its turn: the parallel sub-step decides it
        done = ser_end;
        if (done < n_eval) {
          F = sh.last_open;
          if (valid && tid >= done && !force_par) classify();
        }
        ++n_ser;
        continue;
      }
      force_par = false;
      if (timing && tid == 0) sh.tlast = gtimer();

      // a slot whose seed is not a grower at its turn is void (the state only gets more taken: it never will be)
      if (act && slot >= 0 && !glook) A.doom[i] = 1;
      // the first unfinished seed is a grower with more open assumed-taken points than its thread can look at:
      // the block walks the list (rare)
      if (tid == done && glook && g_ready && g_npend > PEND_CAP && !pend_done) {
        sh.sp_big = 1;
        sh.sp_slot = slot;
      }
      if (act && glook) atomicMin(&sh.first_g, tid);
      // cap the segment so that the hash table stays at most half full
      const bool tiny = act && live && !glook;
      int wsum = tiny ? __popc(want) : 0;
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL_MASK, wsum, o);
        if (lane >= o) wsum += v;
      }
      if (lane == 31) sh.warp_sum[tid >> 5] = wsum;
      __syncthreads();
      if (sh.sp_big) {
        const int g = sh.sp_slot;
        const Slot& sl = S.slots[g];
        const PagedStore st = slot_store(S, g);
        const int64_t n_at = sl.t.n_at;
        bool bad = false;
        for (int64_t k = tid; k < n_at; k += SWEEP_T) {
          const int32_t pt = st.get_at(k);
          if (__ldcg(A.state + 2 * (int64_t)(pt)) == -1) {
            const uint32_t r = __ldcg(A.res + 2 * (int64_t)(pt));
            bad |= r == RES_FREE || !set_has<CSET_BITS>(sh.cset, r);
          }
        }
        if (bad) sh.sp_bad = 1;
        __syncthreads();
        if (tid == done) {
          pend_done = true;
          pend_bad = sh.sp_bad != 0;
        }
      }
      int incl = wsum;  // inclusive prefix of wants in index order
      for (int w = 0; w < (tid >> 5); ++w) incl += sh.warp_sum[w];
      if (tiny && incl > ITEMS) atomicMin(&sh.first_over, tid);
      if (valid) sh_me[tid] = me;
      __syncthreads();
      PHASE(0)
      int seg_hi = sh.first_over < n_eval ? sh.first_over : n_eval;
      if (seg_hi <= done) seg_hi = done + 1;  // a single seed always fits (K-1 <= 31 wants)
      const int first_g = sh.first_g;
      const bool in_seg = act && tid < seg_hi;
      const bool have_g = first_g < seg_hi;  // grower-looking seeds in the segment
      // ---- the wants of the tiny seeds as one flat list in thread order (balanced work for the passes below):
      //      (point, its reservation as gathered, thread) ----
      const int my_off = incl - (tiny ? __popc(want) : 0);
      if (in_seg && tiny && want) {
        int pos = my_off;
#pragma unroll
        for (int j = 1; j < KM; ++j) {
          if (!((want >> j) & 1u))
            continue;
          it_id[pos] = (uint32_t)ids[j];
          it_r[pos] = rs[j];
          it_tid[pos] = (uint16_t)tid;
          ++pos;
        }
      }
      if (tid == seg_hi - 1) sh.n_items = incl;  // wants of the tiny seeds below seg_hi
      if (in_seg && glook) {
        uint32_t h = (me * 2654435761u) >> (32 - GSET_BITS);
        for (;;) {
          const uint32_t old = atomicCAS(sh.gkey + h, 0xffffffffu, me);
          if (old == 0xffffffffu) {
            sh.gval[h] = (uint32_t)tid;
            break;
          }
          h = (h + 1) & (GSET - 1);
        }
      }
      __syncthreads();
      PHASE(1)
      const int n_items = sh.n_items;
      if (tid == 0 && n_items > 0) sh.log_base = atomicAdd(&S.sc[SC_NLOG], (unsigned long long)n_items);  // (read after two more barriers)
      // ---- every want into the hash table: point -> lowest thread of the segment that wants it ----
      for (int k = tid; k < n_items; k += SWEEP_T) {
        const uint32_t id = it_id[k];
        uint32_t h = (id * 2654435761u) >> (32 - HT_BITS);
        for (;;) {
          const uint32_t old = atomicCAS(hkeys + h, 0xffffffffu, id);
          if (old == 0xffffffffu || old == id) {
            atomicMin(hvals + h, (uint32_t)it_tid[k]);
            it_slot[k] = (uint16_t)h;
            break;
          }
          h = (h + 1) & (HT - 1);
        }
      }
      __syncthreads();
      PHASE(2)
      // lowest thread of the segment that wants point p (SWEEP_T: nobody)
      auto wanted_by = [&](uint32_t p) -> uint32_t {
        uint32_t h = (p * 2654435761u) >> (32 - HT_BITS);
        for (;;) {
          const uint32_t k = hkeys[h];
          if (k == 0xffffffffu)
            return (uint32_t)SWEEP_T;
          if (k == p)
            return hvals[h];
          h = (h + 1) & (HT - 1);
        }
      };
      // ---- conflicts with lower seeds of the segment ----
      if (have_g) {  // a wanted point that a lower grower-looking seed of the segment holds
        for (int k = tid; k < n_items; k += SWEEP_T) {
          const uint32_t r = it_r[k];
          const uint32_t t = it_tid[k];
          if (r != RES_FREE && r < sh_me[t] && gset_find(sh.gkey, sh.gval, r) < t) {
            atomicMin(&sh.first_conf, (int)t);
            atomicAdd(&sh.n_conf, 1);
          }
        }
      }
      bool g_bad = false;  // grower-looking: an assumed-taken point is free at its turn
      if (in_seg && live) {
        bool conf = wanted_by(s) < (uint32_t)tid;
        if (have_g && rs_self != RES_FREE && rs_self < me) conf |= gset_find(sh.gkey, sh.gval, rs_self) < (uint32_t)tid;
        if (glook) {
#pragma unroll
          for (int j = 1; j < KM; ++j)
            if ((want >> j) & 1u) {
              conf |= wanted_by((uint32_t)ids[j]) < (uint32_t)tid;
              if ((held_lo >> j) & 1u) conf |= gset_find(sh.gkey, sh.gval, rs[j]) < (uint32_t)tid;
            }
          if (g_ready) {
            if (pend_done) {
              g_bad = pend_bad;
            } else if (g_npend > PEND_CAP) {
              conf |= tid != done;  // the block walks its list when it is the first of a segment
            } else {
              // open assumptions (free when the sweep began): usually the holder recorded at verification has committed
              // a plane in this sweep or a lower tiny seed of the segment marks the point (shared memory answers);
              // global memory only for the others, all loads of a trip in flight together
              for (int k0 = 0; k0 < g_npend; k0 += 8) {
                int2 pe[8];
                uint32_t need = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) pe[u] = k0 + u < g_npend ? S.pend[(size_t)slot * PEND_CAP + k0 + u] : make_int2(-1, -1);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  if (pe[u].x < 0)
                    continue;
                  const uint32_t r = (uint32_t)pe[u].y;
                  if (r != RES_FREE && set_has<CSET_BITS>(sh.cset, r))
                    continue;  // in a plane this sweep committed
                  if (wanted_by((uint32_t)pe[u].x) < (uint32_t)tid)
                    continue;  // a lower tiny seed of this segment marks it
                  need |= 1u << u;
                }
                if (need) {
                  int32_t ps[8];
                  uint32_t pr[8];
#pragma unroll
                  for (int u = 0; u < 8; ++u) {
                    ps[u] = 0;
                    pr[u] = RES_FREE;
                    if ((need >> u) & 1u) {
                      ps[u] = __ldcg(A.state + 2 * (int64_t)(pe[u].x));
                      pr[u] = __ldcg(A.res + 2 * (int64_t)(pe[u].x));
                    }
                  }
#pragma unroll
                  for (int u = 0; u < 8; ++u) {
                    if (!((need >> u) & 1u) || ps[u] != -1)
                      continue;  // marked by a lower seed since
                    const uint32_t r = pr[u];
                    if (r != RES_FREE && set_has<CSET_BITS>(sh.cset, r))
                      continue;
                    if (r != RES_FREE && gset_find(sh.gkey, sh.gval, r) < (uint32_t)tid) conf = true;  // depends on that grower
                    else g_bad = true;
                  }
                }
              }
            }
          }
        }
        if (conf) {
          atomicMin(&sh.first_conf, tid);
          atomicAdd(&sh.n_conf, 1);
        }
      }
      __syncthreads();
      PHASE(3)
      const int seg_end = sh.first_conf < seg_hi ? sh.first_conf : seg_hi;  // > done: the first one has nobody below it
      const int n_conf = sh.n_conf;
      const bool g_here = first_g < seg_end;  // (uniform)
      int first_fail = SWEEP_T;
      if (g_here) {
        // ---- tiny seeds: a higher transaction that holds a point which is about to be marked is void ----
        for (int k = tid; k < n_items; k += SWEEP_T) {
          const uint32_t r = it_r[k];
          const uint32_t t = it_tid[k];
          if ((int)t < seg_end && r != RES_FREE && r > sh_me[t]) doom_now(r);
        }
        __syncthreads();
        // ---- growers decide ----
        if (in_seg && glook && tid < seg_end) {
          bool ok = slot >= 0 && g_ready && !g_bad && !g_doom0;
          if (ok) ok = ((volatile int*)&sh.dset_over)[0] ? ((volatile uint8_t*)A.doom)[i] == 0 : !set_has<DSET_BITS>(sh.dset, me);
          if (ok && g_len > A.th_count && sh.n_cset >= CSET / 2) ok = false;  // (the set is sized for every slot: never)
          if (!ok) atomicMin(&sh.first_fail, tid);
          n_pendsum += (uint32_t)g_npend;
        }
        __syncthreads();
        first_fail = sh.first_fail;
        PHASE(4)
      }
      const bool failed = first_fail < SWEEP_T;  // (only threads below seg_end decide: first_fail < seg_end then)
      const int seg_end2 = failed ? first_fail : seg_end;
      // ---- commit ----
      if (in_seg && glook && tid < seg_end2) {
        Slot& sl = S.slots[slot];
        if (g_len > A.th_count) {  // :199-202
          const unsigned long long off = atomicAdd(&sh.c_off, (unsigned long long)g_len);
          const unsigned long long pl = atomicAdd(&sh.c_pl, 1ull);
          if ((int64_t)(off + g_len) > A.pool_cap - A.n - 2 || (int64_t)pl >= A.planes_cap) {
            A.ctl[CTL_ERR] = 2;
          } else {
            PlaneRec r;
            r.seed = (int32_t)i; r.pad = 0;
            r.off = (int64_t)off; r.len = g_len;
            r.nrm[0] = sl.t.m.mn0; r.nrm[1] = sl.t.m.mn1; r.nrm[2] = sl.t.m.mn2;
            r.ctr[0] = sl.t.m.mc0; r.ctr[1] = sl.t.m.mc1; r.ctr[2] = sl.t.m.mc2; r.pad2 = 0;
            A.planes[pl] = r;
            sl.pool_off = (int64_t)off;
            sl.status = ST_COMMIT;
            atomicAdd(&sh.n_cset, 1);
            set_insert<CSET_BITS>(sh.cset, me);
          }
        } else {  // roll back (:203-209): nothing persists; the reservations are dropped by the next release pass
          sl.status = ST_ROLLED;
        }
        atomicAdd(&sh.c_steps, sl.steps);
        atomicAdd(&sh.c_tx, 1ull);
        ++n_grow;
      }
      if (in_seg && tiny && tid < seg_end2) ++ntiny;
      {
        // orphan marks of the tiny transactions (:233 then :238-239).  Only the owner mark is on the sweeper's path;
        // what the mark means for others (reservation void, alive bit, seed of a waiting grower taken) is logged and
        // applied by a parallel kernel after the sweep.  The list is in thread order: the marks of the committed seeds
        // are a prefix of it.  In a segment without growers the higher holders are doomed here.
        const unsigned long long lbase = sh.log_base;
        for (int k = tid; k < n_items; k += SWEEP_T) {
          const uint32_t id = it_id[k];
          const uint32_t t = it_tid[k];
          const uint32_t seed = sh_me[t];
          const bool commit = (int)t < seg_end2;  // (thread order: the committed seeds' wants are a prefix of the list)
          if (commit) {
            atomicMin(reinterpret_cast<uint32_t*>(A.state) + 2 * (int64_t)id, seed);  // the lower seed owns a shared point
            if (!g_here) {
              const uint32_t r = it_r[k];
              if (r != RES_FREE && r > seed) A.doom[r] = 1;
            }
          }
          const unsigned long long pos = lbase + (unsigned long long)k;
          if (pos < S.marklog_cap) S.marklog[pos] = make_uint2(commit ? id : 0xffffffffu, seed);
          const int h = it_slot[k];  // the table is emptied item by item
          hkeys[h] = 0xffffffffu;
          hvals[h] = 0xffffffffu;
        }
      }
      if (failed && tid == first_fail) {
        // the sweep stops here: still growing (the head), no slot yet (the scout gives it one), or void (released next
        // round, then re-run)
        sh.last_open = i;
        if (slot < 0 || S.slots[slot].status != ST_RUNNING) S.sc[SC_STUCK] = 1;
        if (slot < 0) S.sc[SC_END_NOSLOT] += 1;
        else if (S.slots[slot].status == ST_RUNNING) S.sc[SC_END_RUNNING] += 1;
        else if (S.slots[slot].status == ST_FINISHED && S.slots[slot].verified == 0 && !((volatile uint8_t*)A.doom)[i]) S.sc[SC_END_UNVERIFIED] += 1;
        if (g_ready && g_bad) {
          A.doom[i] = 1;
          atomicAdd(&S.sc[SC_ATFAIL], 1ull);
        }
      }
      if (have_g)
        for (int k = tid; k < GSET; k += SWEEP_T) {
          sh.gkey[k] = 0xffffffffu;
          sh.gval[k] = (uint32_t)SWEEP_T;
        }
      // (no fence: every reader of these marks is in this block, and the barrier orders the block's accesses)
      if (!failed && tid == seg_end2 && valid) sh.last_open = i;  // first seed of the batch that is not finished
      __syncthreads();
      PHASE(5)
      if (!failed && seg_end < seg_hi && n_conf * 8 > seg_hi - done) serial = true;  // conflicts are dense here (scan order)
      done = seg_end2;
      if (failed) {
        stop = true;
        F = sh.last_open;
      } else if (done < n_eval) {
        F = sh.last_open;
        // ---- refresh: the states of what this thread already knows (one load level) ----
        if (valid && tid >= done) classify();
        PHASE(6)
      }
    }
    if (!stop) {  // the whole batch is done
      if (n_live > n_eval) F = sh.last_seed + 1;  // more live seeds in the stretch than threads
      else F = stretch_end;
      if (F > A.n) F = A.n;
    }
    __syncthreads();
  }

  if (tid == 0) {
    *(volatile unsigned long long*)&S.sc[SC_STOP] = 1ull;  // the background slice ends with the sweep
    S.sc[SC_T_FRONT] += t_front;
    for (int k = 0; k < 7; ++k) S.sc[SC_SWEEP_PH + k] += sh.tph[k];
    S.sc[SC_N_SLOW] += n_sub;
    S.sc[SC_N_SER] += n_ser;
    A.ctl[CTL_FRONTIER] = (unsigned long long)F;
    A.ctl[CTL_POOL] = sh.c_off;
    A.ctl[CTL_PLANES] = sh.c_pl;
    A.ctl[CTL_STEPS] += sh.c_steps;
    A.ctl[CTL_TX] += sh.c_tx;
    S.sc[SC_SWEEP_ITERS] += iters;
    S.sc[SC_SWEEP_NS] += gtimer() - t_begin;
  }
  // tiny transactions committed: one atomic per warp
  for (int o = 16; o; o >>= 1) ntiny += __shfl_down_sync(FULL_MASK, ntiny, o);
  for (int o = 16; o; o >>= 1) n_grow += __shfl_down_sync(FULL_MASK, n_grow, o);
  for (int o = 16; o; o >>= 1) n_pendsum += __shfl_down_sync(FULL_MASK, n_pendsum, o);
  if (lane == 0 && ntiny) {
    atomicAdd(&S.sc[SC_TINY], (unsigned long long)ntiny);
    atomicAdd(&A.ctl[CTL_TX], (unsigned long long)ntiny);
    atomicAdd(&A.ctl[CTL_STEPS], (unsigned long long)ntiny);
  }
  if (lane == 0 && n_grow) atomicAdd(&S.sc[SC_N_GROW], (unsigned long long)n_grow);
  if (lane == 0 && n_pendsum) atomicAdd(&S.sc[SC_T_SLOW], (unsigned long long)n_pendsum);
}

// ---- the sweeper's L2 warmer -------------------------------------------------------------------------------------
// The sweeper is one block bound by three dependent load levels per batch (seed -> row -> owner / reservation records),
// most of them DRAM misses.  This block runs beside it on another SM, a few stretches AHEAD of the frontier the sweeper
// publishes, and walks the same chain for every live seed -- inv / gmask / slot, the depth-0 columns of the row, then a
// prefetch.global.L2 of every record the sweeper is going to gather (and of the slot, its model and its open
// assumptions for a seed that has one).  It writes nothing: the sweeper finds its lines in L2 instead of in DRAM.
// Ends with the sweep (SC_STOP), when it reaches the end of the cloud, or when no sweeper shows up.
constexpr int PF_T = 512;
__global__ void __launch_bounds__(PF_T) spec_sweep_prefetch_kernel(SpecArgs S, int64_t lead)
{
  const GrowArgs& A = S.A;
  __shared__ unsigned long long sh_f;
  __shared__ int sh_stop;
  const int tid = threadIdx.x;
  const int K = A.K;
  const int64_t n_words = (A.n + 31) >> 5;
  const int64_t F0 = (int64_t)A.ctl[CTL_FRONTIER];
  int64_t P = F0 & ~31LL;  // first seed of the next stretch to warm
  const unsigned long long t0 = gtimer();
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const unsigned long long now = gtimer();
      const unsigned long long on = *(volatile unsigned long long*)&S.sc[SC_SWEEP_ON];
      int stop = *(volatile unsigned long long*)&S.sc[SC_STOP] != 0;
      if (!on && now - t0 > 1000000ull) stop = 1;      // no sweeper within a millisecond
      if (now - t0 > 4000000000ull) stop = 1;          // (safety net)
      sh_stop = stop;
      sh_f = *(volatile unsigned long long*)&S.sc[SC_LIVE_F];
    }
    __syncthreads();
    if (sh_stop)
      break;
    int64_t F = (int64_t)sh_f;
    if (F < F0) F = F0;
    if (P < (F & ~31LL)) P = F & ~31LL;  // the sweeper caught up
    if (P >= A.n)
      break;
    if (P >= F + lead) {
      __nanosleep(1000);
      continue;
    }
    // one stretch of the alive bitmap: 4 threads per word, 8 seeds each
    const int64_t wi = (P >> 5) + (tid >> 2);
    uint32_t w = wi < n_words ? __ldcg(S.alive + wi) : 0u;
    w = (w >> (8 * (tid & 3))) & 0xffu;
    const int64_t i0 = (wi << 5) + 8 * (tid & 3);
    while (w) {
      const int b = __ffs(w) - 1;
      w &= w - 1;
      const int64_t i = i0 + b;
      if (i < F || i >= A.n)
        continue;
      const uint32_t s = __ldg(A.inv + i);
      const uint32_t m = __ldg(S.gmask + i);
      const int32_t slot = __ldcg(A.slotof + i);
      prefetch_l2(A.state + 2 * (int64_t)s);
      const int32_t* row = A.nbr + (int64_t)s * K;
      uint32_t mm = m;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const int32_t id = __ldg(row + j);
        if (id >= 0) prefetch_l2(A.state + 2 * (int64_t)id);
      }
      if (slot >= 0) {
        const char* sl = reinterpret_cast<const char*>(&S.slots[slot]);
        prefetch_l2(sl);
        prefetch_l2(sl + 128);
        prefetch_l2(A.doom + i);
        // the open assumptions of a finished slot: the records the sweeper asks for when it decides the grower
        int np = __ldcg(&S.slots[slot].n_pend);
        np = np < 0 ? 0 : (np > PEND_CAP ? PEND_CAP : np);
        for (int k = 0; k < np; ++k) prefetch_l2(A.state + 2 * (int64_t)__ldcg(&S.pend[(size_t)slot * PEND_CAP + k].x));
      }
    }
    P += 32LL * (PF_T / 4);
  }
}

__global__ void __launch_bounds__(TPB) spec_alive_init_kernel(uint32_t* alive, int64_t n, int64_t n_alloc_words)
{
  const int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (k >= n_alloc_words)
    return;
  const int64_t lo = k * 32;
  alive[k] = lo + 32 <= n ? 0xffffffffu : (lo >= n ? 0u : (0xffffffffu >> (32 - (int)(n - lo))));
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
    for (int k = 0; k < 16; ++k) S.sc[k] = 0;
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
  // window of seeds the scout looks at / grower slots (tunables: BSEG_WINDOW, BSEG_SLOTS)
  const int64_t CMAX = getenv("BSEG_WINDOW") ? atoll(getenv("BSEG_WINDOW")) : (1 << 18);
  SpecArgs S;
  S.G = getenv("BSEG_SLOTS") ? atoi(getenv("BSEG_SLOTS")) : 1024;
  // page pool: room for every point once plus one page per slot, capped at 64 Mi entries
  int64_t pages = (3 * n) / PAGE_SIZE + 2 * S.G;
  if (pages > 32768) pages = 32768;
  S.n_pool_pages = (uint32_t)pages;
  if (S.G > CSET / 2) S.G = CSET / 2;  // every plane a sweep accepts names its seed in the sweeper's set
  const size_t slot_bytes = (size_t)S.G * sizeof(Slot) + (size_t)S.G * 8 + 256;
  const size_t pend_bytes = (size_t)S.G * PEND_CAP * 8 + 16;
  RC_CHECK(dev_ensure(c, c->g_tx, slot_bytes + (size_t)S.G * MAX_PAGES_PER_SLOT * 4 + (size_t)pages * 4 + pend_bytes + 64));
  // flag[CMAX+4] | bsum[blocks+4] | gmask[n] | slotof[n] | atby[n] | alive[words] | doom[n] | hinted[n]
  const int64_t alive_words = (n + 31) / 32 + SWEEP_WORDS + 4;
  const int64_t n_bsum = ceil_div64(CMAX, TPB) + 4;
  RC_CHECK(dev_ensure(c, c->g_spec, (size_t)(CMAX + 4 + n_bsum) * 4 + (size_t)n * 14 + (size_t)alive_words * 4 + 256));
  RC_CHECK(dev_ensure(c, c->g_queue, (size_t)pages * PAGE_SIZE * 16 + 256));
  RC_CHECK(dev_ensure(c, c->g_marklog, ((size_t)n + 4096) * sizeof(uint2)));
  S.slots = dptr<Slot>(c->g_tx);
  S.free_ids = reinterpret_cast<uint32_t*>(S.slots + S.G);
  S.ptabs = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(c->g_tx.p) + slot_bytes);
  S.pool.free_pages = S.ptabs + (size_t)S.G * MAX_PAGES_PER_SLOT;
  S.pend = reinterpret_cast<int2*>((reinterpret_cast<uintptr_t>(S.pool.free_pages + pages) + 15) & ~(uintptr_t)15);
  S.pool.stack_pages = dptr<int2>(c->g_queue);
  S.pool.list_pages = reinterpret_cast<int32_t*>(S.pool.stack_pages + (size_t)pages * PAGE_SIZE);
  S.pool.at_pages = S.pool.list_pages + (size_t)pages * PAGE_SIZE;
  S.flag = dptr<uint32_t>(c->g_spec);
  S.bsum = S.flag + CMAX + 4;
  uint32_t* gmask = S.bsum + n_bsum;
  S.gmask = gmask;
  A.slotof = reinterpret_cast<int32_t*>(gmask + n);
  A.atby = reinterpret_cast<uint32_t*>(A.slotof + n);
  S.alive = A.atby + n;
  A.doom = reinterpret_cast<uint8_t*>(S.alive + alive_words);
  S.hinted = A.doom + n;
  S.marklog = dptr<uint2>(c->g_marklog);
  S.marklog_cap = (unsigned long long)n + 4096;
  S.sc = A.ctl + 8;
  S.pool.n_free = &S.sc[SC_POOLFREE];
  A.stop_flag = &S.sc[SC_STOP];
  A.slice_min_ns = getenv("BSEG_SLICE_MIN_US") ? 1000ull * strtoull(getenv("BSEG_SLICE_MIN_US"), nullptr, 10) : 0ull;
  A.frontier = 0;
  S.A = A;
  CU_CHECK(c, cudaMemsetAsync(A.slotof, 0xff, (size_t)n * 8, c->stream));  // slotof = -1, atby = none
  CU_CHECK(c, cudaMemsetAsync(A.doom, 0, (size_t)n * 2, c->stream));
  {
    const int64_t m = pages > S.G ? pages : S.G;
    spec_init_kernel<<<(unsigned)ceil_div64(m, TPB), TPB, 0, c->stream>>>(S);
    KLAUNCH_CHECK(c);
    gmask_kernel<<<(unsigned)ceil_div64(n, TPB), TPB, 0, c->stream>>>(A, gmask);
    KLAUNCH_CHECK(c);
    spec_alive_init_kernel<<<(unsigned)ceil_div64(alive_words, TPB), TPB, 0, c->stream>>>(S.alive, n, alive_words);
    KLAUNCH_CHECK(c);
  }
  if (!c->attr_sweep_set) {
    CU_CHECK(c, cudaFuncSetAttribute(spec_sweep_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<16>::SMEM));
    CU_CHECK(c, cudaFuncSetAttribute(spec_sweep_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<32>::SMEM));
    c->attr_sweep_set = true;
  }

  int64_t F = 0, rounds = 0, stalls = 0, fallbacks = 0;
  const bool dbg = getenv("BSEG_DEBUG") != nullptr;
  // background slices beside the sweeper (BSEG_BG=0: off); bounded in warp iterations even if the stop flag never came
  const bool bg_on = !(getenv("BSEG_BG") && atoi(getenv("BSEG_BG")) == 0);
  const unsigned long long bg_budget = 1ull << 20;
  // the sweeper's L2 warmer (BSEG_PF=0: off; BSEG_PF_LEAD: how many seeds it may run ahead of the frontier)
  const bool pf_on = bg_on && !(getenv("BSEG_PF") && atoi(getenv("BSEG_PF")) == 0);
  const int64_t pf_lead = getenv("BSEG_PF_LEAD") ? atoll(getenv("BSEG_PF_LEAD")) : 32768;
  // the host runs one round ahead of the control block it has seen (BSEG_PIPE=0: launch, wait, look, launch)
  const bool pipe_on = !(getenv("BSEG_PIPE") && atoi(getenv("BSEG_PIPE")) == 0);
  if (bg_on && !c->grow_hi) {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // (numerically lower = higher priority)
    CU_CHECK(c, cudaStreamCreateWithPriority(&c->grow_hi, cudaStreamNonBlocking, prio_hi));
    CU_CHECK(c, cudaStreamCreateWithPriority(&c->grow_lo, cudaStreamNonBlocking, prio_lo));
    CU_CHECK(c, cudaStreamCreateWithPriority(&c->grow_pf, cudaStreamNonBlocking, prio_lo));
    CU_CHECK(c, cudaEventCreateWithFlags(&c->grow_fork, cudaEventDisableTiming));
    for (int k = 0; k < 3; ++k) CU_CHECK(c, cudaEventCreateWithFlags(&c->grow_join[k], cudaEventDisableTiming));
  }
  float pt[4] = {0, 0, 0, 0};
  for (int k = 0; k < 12; ++k)  // two sets of phase events + "control block copied" (a round in flight, a round being looked at)
    if (!c->grow_ev[k]) CU_CHECK(c, cudaEventCreate(&c->grow_ev[k]));
  if (!c->grow_pinned) CU_CHECK(c, cudaHostAlloc(&c->grow_pinned, 2 * 64 * sizeof(unsigned long long), cudaHostAllocDefault));
  unsigned long long ctl[64] = {0};
  const unsigned sb = (unsigned)((S.G + GW - 1) / GW);
  // slice length in warp iterations of the head (two-node engine: ~1.6 calls each, a skip batch weighs 3)
  const unsigned long long budget = getenv("BSEG_SLICE") ? strtoull(getenv("BSEG_SLICE"), nullptr, 10) : 2560;
  // every kernel of a round reads the frontier from the control block: a round can be launched before the host has seen
  // the result of the one before it.  Past the end of the cloud a round is a no-op.
  const int64_t CW = n < CMAX ? n : CMAX;
  const unsigned gb = (unsigned)ceil_div64(CW, TPB);
  S.A.frontier = -1;
  int64_t enq = 0;  // rounds launched
  auto enqueue_round = [&](int set) -> int {
    cudaEvent_t* pe = c->grow_ev + 6 * set;
    cudaEventRecord(pe[0], c->stream);
    if (enq > 0) {
      spec_mark_release_kernel<<<(S.G + TPB - 1) / TPB, TPB, 0, c->stream>>>(S);
      KLAUNCH_CHECK(c);
      spec_release_entries_kernel<<<dim3(RCH, S.G), TPB, 0, c->stream>>>(S);
      KLAUNCH_CHECK(c);
      spec_release_slots_kernel<<<(S.G + TPB - 1) / TPB, TPB, 0, c->stream>>>(S);
      KLAUNCH_CHECK(c);
      cudaEventRecord(pe[1], c->stream);
      spec_scout_kernel<<<gb, TPB, 0, c->stream>>>(S, CW);
      KLAUNCH_CHECK(c);
      spec_assign_kernel<<<gb, TPB, 0, c->stream>>>(S, CW);
      KLAUNCH_CHECK(c);
      spec_pop_free_kernel<<<1, 1, 0, c->stream>>>(S);
      KLAUNCH_CHECK(c);
      cudaEventRecord(pe[2], c->stream);
      spec_grow_kernel<<<sb, GW * 32, 0, c->stream>>>(S, budget, 0);
      KLAUNCH_CHECK(c);
      spec_preverify_kernel<<<dim3(RCH, S.G), TPB, 0, c->stream>>>(S);
      KLAUNCH_CHECK(c);
    } else {
      cudaEventRecord(pe[1], c->stream);
      cudaEventRecord(pe[2], c->stream);
    }
    cudaEventRecord(pe[3], c->stream);
    if (bg_on && enq > 0) {
      // the sweeper (one block, a whole SM: launched first, on the high-priority stream) and the background slice of
      // the growers side by side; both rejoin c->stream before the side effects of the sweep are applied
      cudaEventRecord(c->grow_fork, c->stream);
      cudaStreamWaitEvent(c->grow_hi, c->grow_fork, 0);
      cudaStreamWaitEvent(c->grow_lo, c->grow_fork, 0);
      if (A.K <= 16) spec_sweep_kernel<16><<<1, SWEEP_T, SweepCfg<16>::SMEM, c->grow_hi>>>(S);
      else spec_sweep_kernel<32><<<1, SWEEP_T, SweepCfg<32>::SMEM, c->grow_hi>>>(S);
      KLAUNCH_CHECK(c);
      if (pf_on) {
        cudaStreamWaitEvent(c->grow_pf, c->grow_fork, 0);
        spec_sweep_prefetch_kernel<<<1, PF_T, 0, c->grow_pf>>>(S, pf_lead);
        KLAUNCH_CHECK(c);
        cudaEventRecord(c->grow_join[2], c->grow_pf);
      }
      SpecArgs Sbg = S;
      Sbg.A.flags |= GF_BG;
      spec_grow_kernel<<<sb, GW * 32, 0, c->grow_lo>>>(Sbg, bg_budget, 1);
      KLAUNCH_CHECK(c);
      cudaEventRecord(c->grow_join[0], c->grow_hi);
      cudaEventRecord(c->grow_join[1], c->grow_lo);
      cudaStreamWaitEvent(c->stream, c->grow_join[0], 0);
      cudaStreamWaitEvent(c->stream, c->grow_join[1], 0);
      if (pf_on) cudaStreamWaitEvent(c->stream, c->grow_join[2], 0);
    } else {
      // (the first sweep runs before anything is in flight: it stops at the first grower)
      if (A.K <= 16) spec_sweep_kernel<16><<<1, SWEEP_T, SweepCfg<16>::SMEM, c->stream>>>(S);
      else spec_sweep_kernel<32><<<1, SWEEP_T, SweepCfg<32>::SMEM, c->stream>>>(S);
      KLAUNCH_CHECK(c);
    }
    spec_apply_marks_kernel<<<c->num_sms * 4, TPB, 0, c->stream>>>(S);
    KLAUNCH_CHECK(c);
    spec_apply_commits_kernel<<<dim3(RCH, S.G), TPB, 0, c->stream>>>(S);  // (also empties the mark log)
    KLAUNCH_CHECK(c);
    cudaEventRecord(pe[4], c->stream);
    CU_CHECK(c, cudaMemcpyAsync(c->grow_pinned + 64 * set, A.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, c->stream));
    cudaEventRecord(pe[5], c->stream);
    ++enq;
    return 0;
  };
  // wait for the round launched with event set `set`: its control block and phase times
  auto collect_round = [&](int set) -> int {
    cudaEvent_t* pe = c->grow_ev + 6 * set;
    CU_CHECK(c, cudaEventSynchronize(pe[5]));
    memcpy(ctl, c->grow_pinned + 64 * set, sizeof(ctl));
    ++rounds;
    if (rounds > 1)
      for (int k = 0; k < 4; ++k) {
        float ms = 0;
        cudaEventElapsedTime(&ms, pe[k], pe[k + 1]);
        pt[k] += ms;
      }
    return 0;
  };
  // what the control block of a finished round says.  0: go on; 1: done (or a capacity error: ctl[CTL_ERR]); 2: the head
  // cannot get a slot (no pages left, or a plane too large for the page table): grow it alone in the flat region
  auto look = [&]() -> int {
    if (ctl[CTL_ERR])
      return 1;
    const int64_t Fn = (int64_t)ctl[CTL_FRONTIER];
    if (Fn < F)
      return bseg_fail(c, BSEG_E_STATE, "plane grower: frontier moved backwards");
    const bool head_has_slot = ctl[8 + SC_HEAD_SLOT] != 0;
    const bool stuck = Fn == F && rounds > 1 && !head_has_slot;
    F = Fn;
    if (F >= n)
      return 1;
    if (rounds > 8 * n + 1024)
      return bseg_fail(c, BSEG_E_STATE, "plane grower: no progress");
    if (stuck) {
      if (++stalls >= 2)
        return 2;
    } else {
      stalls = 0;
    }
    return 0;
  };
  int in_flight = 0, oldest = 0;  // event sets alternate: `oldest` is the one to collect next
  for (;;) {
    while (in_flight < (pipe_on ? 2 : 1)) {
      RC_CHECK(enqueue_round((oldest + in_flight) & 1));
      ++in_flight;
    }
    RC_CHECK(collect_round(oldest));
    oldest ^= 1;
    --in_flight;
    int r = look();
    if (r < 0) {
      cudaStreamSynchronize(c->stream);
      return r;
    }
    if (r != 0) {
      // nothing may be in flight when the loop ends or when the sequential engine takes the head
      while (in_flight > 0) {
        RC_CHECK(collect_round(oldest));
        oldest ^= 1;
        --in_flight;
        if (r == 2) {
          const int r2 = look();
          if (r2 < 0)
            return r2;
          if (r2 == 1) r = 1;
          else if (r2 == 0) r = 0;  // it got a slot after all
        }
      }
      if (r == 1)
        break;
      if (r == 2) {
        launch_grow_seq(c, S.A, true, 1, 1ull << 40, 1);
        KLAUNCH_CHECK(c);
        RC_CHECK(read_back(c, ctl, A.ctl, sizeof(ctl)));
        if (ctl[CTL_ERR])
          break;
        stalls = 0;
        ++fallbacks;
        F = (int64_t)ctl[CTL_FRONTIER];
        if (F >= n)
          break;
      }
    }
  }
  if (dbg) {
    fprintf(stderr, "[bseg] head: %llu calls in %llu warp steps\n", ctl[8 + SC_HEAD_STEPS], ctl[8 + SC_HEAD_ITERS]);
    const unsigned long long* sd = ctl + 8 + SC_SKIP_DBG;
    if (sd[3])
      fprintf(stderr, "[bseg] head regular steps %llu (%llu accepting): cycles/step want %llu, reserve+ballots %llu, dfs %llu, row %llu, accumulate %llu, "
              "gather issue+update %llu, race+tail %llu, between %llu\n", sd[3], sd[16], sd[8] / sd[3], sd[9] / sd[3], sd[10] / sd[3], sd[11] / sd[3],
              sd[12] / sd[3], sd[13] / sd[3], sd[14] / sd[3], sd[15] / sd[3]);
    if (sd[0])
      fprintf(stderr, "[bseg] head skip batches %llu: cycles/batch enumerate %llu, evaluate %llu; pairs/batch %.1f; regular steps %llu, "
              "batch time %.1f ms of head %.1f ms\n", sd[0], sd[1] / sd[0], sd[2] / sd[0], (double)sd[4] / (double)sd[0], sd[3],
              (double)(sd[1] + sd[2]) / 1.965e6, ctl[8 + SC_HEAD_NS] / 1e6);
    fprintf(stderr, "[bseg] sweeps ended at a grower without a slot %llu, still growing %llu, finished beside the sweeper (unverified) %llu, void or failed %lld\n",
            ctl[8 + SC_END_NOSLOT], ctl[8 + SC_END_RUNNING], ctl[8 + SC_END_UNVERIFIED],
            (long long)rounds - (long long)(ctl[8 + SC_END_NOSLOT] + ctl[8 + SC_END_RUNNING] + ctl[8 + SC_END_UNVERIFIED]));
    fprintf(stderr, "[bseg] rounds %lld: release %.1f ms, scout+assign %.1f ms, slices %.1f ms, sweep+apply %.1f ms; background slices %s, %llu steps\n",
            (long long)rounds, pt[0], pt[1], pt[2], pt[3], bg_on ? "on" : "off", ctl[8 + SC_BG_STEPS]);
  }
  c->tm.grow_slice_ms = pt[2];
  c->tm.grow_sweep_ms = pt[3];
  if (dbg)
    fprintf(stderr, "[bseg] sweeper: front %.1f ms of %.1f ms (%llu segments, %llu of them serial stretches, %llu growers decided with %llu open assumed-taken points), %llu batches\n",
            ctl[8 + SC_T_FRONT] / 1e6, ctl[8 + SC_SWEEP_NS] / 1e6, ctl[8 + SC_N_SLOW], ctl[8 + SC_N_SER], ctl[8 + SC_N_GROW], ctl[8 + SC_T_SLOW], ctl[8 + SC_SWEEP_ITERS]);
  if (dbg)
    fprintf(stderr, "[bseg] sweeper sub-step phases (ms): classify-flags+prefix %.1f, item list %.1f, insert %.1f, conflicts %.1f, dooms+decide %.1f, commit %.1f, refresh %.1f\n",
            ctl[8 + SC_SWEEP_PH] / 1e6, ctl[8 + SC_SWEEP_PH + 1] / 1e6, ctl[8 + SC_SWEEP_PH + 2] / 1e6, ctl[8 + SC_SWEEP_PH + 3] / 1e6,
            ctl[8 + SC_SWEEP_PH + 4] / 1e6, ctl[8 + SC_SWEEP_PH + 5] / 1e6, ctl[8 + SC_SWEEP_PH + 6] / 1e6);
  c->tm.grow_rounds = rounds;
  c->tm.grow_wasted_steps = (int64_t)ctl[8 + SC_WASTED];
  c->tm.grow_sweep_iters = (int64_t)ctl[8 + SC_SWEEP_ITERS];
  c->tm.grow_tiny_tx = (int64_t)ctl[8 + SC_TINY];
  c->tm.grow_seq_fallbacks = fallbacks;
  c->tm.grow_at_fails = (int64_t)ctl[8 + SC_ATFAIL];
  c->tm.grow_head_steps = (int64_t)ctl[8 + SC_HEAD_STEPS];
  c->tm.grow_head_ns = (int64_t)ctl[8 + SC_HEAD_NS];
  c->tm.grow_sweep_ns = (int64_t)ctl[8 + SC_SWEEP_NS];
  return 0;
}
