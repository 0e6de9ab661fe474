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
// of Broad steps for every running slot (ends when the head slot finishes) -> sweep.  Plane ids are ordinal
// in seed order and are assigned after the fact (grow.cu finalize).
#include <cstdlib>

#include "grow.cuh"

namespace {

constexpr int TPB = 256;
constexpr int GW = 4;          // warps per block in slot kernels
constexpr int SWEEP_T = 1024;  // seeds per sweeper batch == threads of the sweeper block
constexpr int SWEEP_WORDS = 128;  // bitmap words (4096 seeds) scanned per batch for up to SWEEP_T live seeds
constexpr int PT_CACHE = 256;      // sweeper: cached page-table entries of the slot being committed
constexpr int HT = 16384;      // sweeper hash table slots (keys + vals = 128 KB of shared memory)

enum { ST_FREE = 0, ST_RUNNING = 1, ST_FINISHED = 2, ST_DEAD = 3, ST_RELEASING = 4, ST_DONE = 5 };  // DONE: committed by the sweeper, slot and pages still to be returned
// spec control block (A.ctl + 8)
enum {
  SC_NFREE = 0,       // free slots
  SC_NCAND = 1,       // candidates of the window (scan total)
  SC_STOP = 2,        // head finished: slices end
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
  const uint32_t* gmask;  // [n] ORIGINAL index space (the sweeper and the scout walk seeds in index order)
  uint32_t* flag;   // [C] candidate flags -> exclusive scan
  uint32_t* free_ids;
  uint8_t* hinted;    // [n] original index space: this tiny transaction has published hints
  uint32_t* alive;    // [ceil(n/32)] original index space: bit = the point may still be free (filter)
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
constexpr int RCH = 16;
// which slots go: decided ONCE (releasing a slot can doom others through the early notification, and the
// three kernels must agree)
__global__ void __launch_bounds__(TPB) spec_mark_release_kernel(SpecArgs S)
{
  const int g = blockIdx.x * TPB + threadIdx.x;
  if (g >= S.G)
    return;
  Slot& sl = S.slots[g];
  if (sl.status == ST_FREE || sl.status == ST_DONE)
    return;
  if (sl.status == ST_DEAD || S.A.doom[sl.seed_i]) sl.status = ST_RELEASING;
}

__global__ void __launch_bounds__(TPB) spec_release_entries_kernel(SpecArgs S)
{
  const GrowArgs& A = S.A;
  const int g = blockIdx.y;
  const Slot& sl = S.slots[g];
  if (sl.status != ST_RELEASING)
    return;
  const int32_t i = sl.seed_i;
  const PagedStore st = slot_store(S, g);
  const int64_t len = sl.started ? sl.t.len : 0;
  for (int64_t e = 1 + (int64_t)blockIdx.x * TPB + threadIdx.x; e < len; e += (int64_t)RCH * TPB) {
    const int32_t pt = st.get(e);
    if (atomicCAS(A.res + pt, (uint32_t)i, RES_FREE) == (uint32_t)i) unreserve_notify(A.atby, A.slotof, A.doom, pt);
  }
}

__global__ void __launch_bounds__(TPB) spec_release_slots_kernel(SpecArgs S)
{
  const int g = blockIdx.x * TPB + threadIdx.x;
  if (g >= S.G)
    return;
  Slot& sl = S.slots[g];
  if (sl.status != ST_RELEASING && sl.status != ST_DONE)
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
    const uint32_t r = __ldcg(A.res + e.x);
    if (r != RES_FREE) {
      if (r != e.y) A.doom[r] = 1;  // a grower ahead of the sweeper held it: void
      A.res[e.x] = RES_FREE;        // (a hint of this or a higher tiny transaction: obsolete)
    }
    const int32_t w = __ldg(&A.pts[e.x].w);
    atomicAnd(S.alive + (w >> 5), ~(1u << (w & 31)));
    if (__ldcg(A.slotof + w) >= 0) A.doom[w] = 1;  // the grower seeded at this point lost its seed
  }
}

// alive bits of the points of the planes this sweep committed: pool entries [SC_POOL0, CTL_POOL)
__global__ void __launch_bounds__(TPB) spec_apply_planes_kernel(SpecArgs S)
{
  const GrowArgs& A = S.A;
  const unsigned long long lo = S.sc[SC_POOL0], hi = A.ctl[CTL_POOL];
  for (unsigned long long k = lo + (unsigned long long)blockIdx.x * TPB + threadIdx.x; k < hi;
       k += (unsigned long long)gridDim.x * TPB) {
    const int32_t w = __ldg(&A.pts[A.pool[k]].w);
    atomicAnd(S.alive + (w >> 5), ~(1u << (w & 31)));
  }
}

__global__ void spec_reset_log_kernel(SpecArgs S)
{
  S.sc[SC_NLOG] = 0;
  S.sc[SC_POOL0] = S.A.ctl[CTL_POOL];
}

// ---- K1: scout -- the window [F, F+C) ahead of the sweeper -------------------------------------------------
// Every seed of the window without a slot is evaluated against the committed state plus the reservations
// of LOWER in-flight transactions (reservations below F are stale and count as free):
//   dead (seed taken, or reserved by a lower transaction): retract the hints it may have published;
//   grower-looking (all K-1 neighbours pass depth 0, free and unreserved by anything lower): candidate;
//   tiny-looking: publish the marks it is going to make, atomicMin(res[pt], i) -- a higher grower that
//     held the point is doomed.
__device__ __forceinline__ uint32_t eff_res(uint32_t r, uint32_t F) { return r < F ? RES_FREE : r; }

__global__ void __launch_bounds__(TPB) spec_scout_kernel(SpecArgs S, int64_t F, int64_t C)
{
  const GrowArgs& A = S.A;
  const int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (t == 0) {
    S.sc[SC_STOP] = 0;
    S.sc[SC_NCAND] = 0;
    S.sc[SC_NASSIGN] = 0;
  }
  if (t >= C)
    return;
  const int64_t i = F + t;
  const uint32_t me = (uint32_t)i, fr = (uint32_t)F;
  uint32_t cand = 0;
  if (__ldcg(A.slotof + i) < 0) {
    const uint32_t s = __ldg(A.inv + i);
    const int K = A.K;
    const uint32_t m = __ldg(S.gmask + i);
    const int32_t* row = A.nbr + (int64_t)s * K;
    const bool hinted = S.hinted[i] != 0;
    const bool dead = __ldcg(A.state + s) != -1 || eff_res(__ldcg(A.res + s), fr) < me;
    uint32_t want = 0;
    bool lower = false;
    if (!dead) {
      uint32_t mm = m;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const int32_t id = __ldg(row + j);
        if (__ldcg(A.state + id) == -1) {
          want |= 1u << j;
          lower |= eff_res(__ldcg(A.res + id), fr) < me;
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
          if (atomicCAS(A.res + pt, me, RES_FREE) == me) unreserve_notify(A.atby, A.slotof, A.doom, pt);
        }
        S.hinted[i] = 0;
      }
    } else if (want) {
      uint32_t mm = want;
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1;
        const int32_t id = __ldg(row + j);
        uint32_t cur = __ldcg(A.res + id);
        for (;;) {
          if (cur != RES_FREE && cur >= fr && cur <= me)
            break;  // ours already, or a lower transaction's
          uint32_t prev;
          if (cur < fr) {  // stale: replace
            prev = atomicCAS(A.res + id, cur, me);
            if (prev != cur) {
              cur = prev;
              continue;
            }
          } else {
            prev = atomicMin(A.res + id, me);
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
  S.flag[t] = cand;
}

// flag[] holds the exclusive scan of the candidate flags; the lowest candidates take the free slots
__global__ void __launch_bounds__(TPB) spec_assign_kernel(SpecArgs S, int64_t F, int64_t C)
{
  const int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (t >= C)
    return;
  const uint32_t r = S.flag[t];
  const uint32_t next = t + 1 < C ? S.flag[t + 1] : (uint32_t)S.sc[SC_NCAND];
  if (next == r)
    return;  // not a candidate
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
  S.A.slotof[F + t] = (int32_t)g;
  S.A.doom[F + t] = 0;
}

__global__ void spec_pop_free_kernel(SpecArgs S, int64_t F)
{
  S.sc[SC_NFREE] -= S.sc[SC_NASSIGN];
  const int32_t g = F < S.A.n ? S.A.slotof[F] : -1;
  S.sc[SC_HEAD_SLOT] = (unsigned long long)(g + 1);
}

// ---- K2: one warp per slot, a time slice of Broad steps with reservations ----------------------------------
__global__ void __launch_bounds__(GW * 32) spec_grow_kernel(SpecArgs S, unsigned long long budget)
{
  const GrowArgs& A = S.A;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * GW + (threadIdx.x >> 5);
  if (g >= S.G)
    return;
  Slot& sl = S.slots[g];
  if (sl.status != ST_RUNNING)
    return;
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
  const bool is_head = (unsigned long long)(g + 1) == S.sc[SC_HEAD_SLOT];
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
  if (lane == 0) {
    sl.t = t;
    sl.started = 1;
    sl.steps += steps;
    // a depth-0 failure under assumptions is a tiny transaction: the sweeper's job
    sl.status = out == TX_RUNNING ? ST_RUNNING : (out == TX_FINISHED ? ST_FINISHED : ST_DEAD);
    // the head sets the pace: when it finishes OR has used its budget the slice is over for everybody
    // (a slot that keeps running after the head has stopped only delays the sweeper)
    if (is_head) *(volatile unsigned long long*)&S.sc[SC_STOP] = 1ull;
  }
}

// ---- K3: the sweeper ------------------------------------------------------------------------------------------
struct SweepShared {
  int first_special;  // lowest thread whose seed needs the slow path (or lies beyond the cloud)
  int first_over;     // lowest thread beyond the hash-table budget
  int first_conf;     // lowest thread whose seed a lower seed of the batch wants
  int stop;
  int sp_slot, sp_np, sp_bad;
  unsigned long long c_off, c_pl;
  int warp_sum[32];
  int n_live;
  int64_t last_seed, last_open, sp_seed;
  uint32_t ptc[PT_CACHE];  // page table of the slot being committed
  uint32_t words[SWEEP_WORDS];
  int pref[SWEEP_WORDS];
};

__device__ __forceinline__ uint32_t sweep_hash(uint32_t p) { return (p * 2654435761u) >> 18; }  // 14 bits

// the sequential authority: walks the seeds from the frontier in index order
template <int KM>
__global__ void __launch_bounds__(SWEEP_T) spec_sweep_kernel(SpecArgs S)
{
  extern __shared__ uint32_t smem_u32[];
  uint32_t* hkeys = smem_u32;        // [HT] point, 0xffffffff = empty
  uint32_t* hvals = smem_u32 + HT;   // [HT] lowest thread of the batch that wants it
  __shared__ SweepShared sh;
  const GrowArgs& A = S.A;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int K = A.K;
  int64_t F = (int64_t)A.ctl[CTL_FRONTIER];
  const int64_t n_words = (A.n + 31) >> 5;
  unsigned long long iters = 0, ntiny = 0;
  const unsigned long long t_begin = gtimer();
  for (int k = tid; k < HT; k += SWEEP_T) {
    hkeys[k] = 0xffffffffu;
    hvals[k] = 0xffffffffu;
  }
  if (tid == 0) sh.stop = 0;
  __syncthreads();
  unsigned long long t_front = 0, t_slow = 0, t_fast = 0, n_slow = 0;
  bool stop = false;

  while (F < A.n && !stop) {
    ++iters;
    const unsigned long long ti0 = gtimer();
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
    uint32_t s = 0, want = 0;
    int32_t slot = -1;
    bool live = false, grower = false;
    int32_t ids[KM];  // neighbour columns that pass depth 0 (registers: every loop over them is unrolled)
#pragma unroll
    for (int j = 1; j < KM; ++j) ids[j] = -1;
    if (valid) {
      // dependent levels of loads: (inv, gmask, slot) -> (state of the seed, row) -> neighbour states
      s = __ldg(A.inv + i);
      const uint32_t m = __ldg(S.gmask + i);
      slot = __ldcg(A.slotof + i);
      const int32_t* row = A.nbr + (int64_t)s * K;
      const int32_t st_s = __ldcg(A.state + s);  // the bitmap is only a filter: this is the truth
#pragma unroll
      for (int j = 1; j < KM; ++j)
        if ((m >> j) & 1u) ids[j] = __ldg(row + j);
      int32_t stj[KM];
#pragma unroll
      for (int j = 1; j < KM; ++j) {
        stj[j] = 0;
        if (ids[j] >= 0) stj[j] = __ldcg(A.state + ids[j]);
      }
      live = st_s == -1;
      if (live) {
#pragma unroll
        for (int j = 1; j < KM; ++j)
          if (ids[j] >= 0 && stj[j] == -1) want |= 1u << j;
        grower = __popc(want) == K - 1;  // :238 -- every neighbour accepted
      }
    }
    t_front += gtimer() - ti0;

    // ---- sub-steps on this batch: threads [0, done) are finished; after every commit the others only
    //      re-gather the states of what they already know (one load level instead of four) ----
    int done = 0;
    while (done < n_eval && !stop) {
      const unsigned long long ti1 = gtimer();
      bool state_changed = true;  // false after a roll-back: nothing the other threads know has moved
      if (tid == 0) {
        sh.first_special = SWEEP_T;
        sh.first_over = SWEEP_T;
        sh.first_conf = SWEEP_T;
      }
      __syncthreads();
      const bool act = valid && tid >= done;
      if (act && grower) atomicMin(&sh.first_special, tid);
      // a slot whose seed is not a grower at its turn is void (the state only gets more taken: it never will be)
      if (act && slot >= 0 && !grower) A.doom[i] = 1;
      __syncthreads();
      int first_special = sh.first_special;
      if (first_special > n_eval) first_special = n_eval;

      if (first_special == done) {
        // ---- slow path: the first unfinished seed is a grower at its turn ----
        ++n_slow;
        if (tid == done) {
          sh.sp_slot = slot;
          sh.sp_bad = 0;
          sh.sp_seed = i;
        }
        __syncthreads();
        const int64_t Fs = sh.sp_seed;  // everything before it is done
        F = Fs;
        const int g = sh.sp_slot;
        if (g < 0) {  // no slot: the scout gives it one
          if (tid == 0) S.sc[SC_STUCK] = 1;
          stop = true;
          break;
        }
        Slot& sl = S.slots[g];
        const int status = sl.status;
        const bool doomed = ((volatile uint8_t*)A.doom)[Fs] != 0;
        if (status != ST_FINISHED || doomed) {
          // still growing (the head), or void: released next round, then the scout re-assigns it
          if (tid == 0 && status != ST_RUNNING) S.sc[SC_STUCK] = 1;
          stop = true;
          break;
        }
        const PagedStore st = slot_store(S, g);
        const int64_t len = sl.t.len, n_at = sl.t.n_at;
        // the slot's page table goes to shared memory: one dependent load level less in every loop below
        if (tid < PT_CACHE && tid < sl.n_pages) sh.ptc[tid] = st.ptab[tid];
        if (tid == 0) {
          sh.c_off = A.ctl[CTL_POOL];
          sh.c_pl = A.ctl[CTL_PLANES];
        }
        __syncthreads();
        auto at = [&](int64_t e) -> size_t {
          const int64_t pg = e >> PAGE_SHIFT;
          const uint32_t page = pg < PT_CACHE ? sh.ptc[pg] : st.ptab[pg];
          return ((size_t)page << PAGE_SHIFT) + (size_t)(e & (PAGE_SIZE - 1));
        };
        // every point it treated as taken (reserved by a lower transaction at the time) must be taken now,
        // and every point it accepted must still be free (a lower tiny transaction may have marked it)
        bool bad = false;
        // (four entries per thread and trip: the loads of a trip are independent, one block has to do it all)
        for (int64_t k0 = tid; k0 < n_at; k0 += 4 * SWEEP_T) {
          int32_t pt[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) pt[u] = k0 + u * SWEEP_T < n_at ? S.pool.at_pages[at(k0 + u * SWEEP_T)] : -1;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (pt[u] >= 0) bad |= __ldcg(A.state + pt[u]) == -1;
        }
        for (int64_t e0 = 1 + tid; e0 < len; e0 += 4 * SWEEP_T) {
          int32_t pt[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) pt[u] = e0 + u * SWEEP_T < len ? S.pool.list_pages[at(e0 + u * SWEEP_T)] : -1;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (pt[u] >= 0) bad |= __ldcg(A.state + pt[u]) != -1;
        }
        if (bad) sh.sp_bad = 1;
        __syncthreads();
        if (sh.sp_bad) {
          if (tid == 0) {
            A.doom[Fs] = 1;
            S.sc[SC_STUCK] = 1;
            atomicAdd(&S.sc[SC_ATFAIL], 1ull);
          }
          stop = true;
          break;
        }
        if (len > A.th_count) {  // :199-202
          const unsigned long long off = sh.c_off, pl = sh.c_pl;
          if ((int64_t)(off + len) > A.pool_cap - A.n - 2 || (int64_t)pl >= A.planes_cap) {
            if (tid == 0) A.ctl[CTL_ERR] = 2;
            stop = true;
            break;
          }
          for (int64_t e0 = tid; e0 < len; e0 += 4 * SWEEP_T) {
            int32_t id[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) id[u] = e0 + u * SWEEP_T < len ? S.pool.list_pages[at(e0 + u * SWEEP_T)] : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (id[u] < 0)
                continue;
              const int64_t e = e0 + u * SWEEP_T;
              A.pool[off + e] = id[u];  // (the alive bits of these points are cleared from the pool by the apply kernel)
              if (e >= 1) {
                A.state[id[u]] = (int32_t)Fs;
                A.res[id[u]] = RES_FREE;
              }
            }
          }
          if (tid == 0) {
            PlaneRec r;
            r.seed = (int32_t)Fs; r.pad = 0;
            r.off = (int64_t)off; r.len = len;
            r.nrm[0] = sl.t.m.mn0; r.nrm[1] = sl.t.m.mn1; r.nrm[2] = sl.t.m.mn2;
            r.ctr[0] = sl.t.m.mc0; r.ctr[1] = sl.t.m.mc1; r.ctr[2] = sl.t.m.mc2; r.pad2 = 0;
            A.planes[pl] = r;
            A.ctl[CTL_POOL] = off + (unsigned long long)len;
            A.ctl[CTL_PLANES] = pl + 1;
          }
        } else {  // roll back (:203-209): nothing persists
          state_changed = false;
          for (int64_t e = 1 + tid; e < len; e += SWEEP_T) {
            const int32_t pt = S.pool.list_pages[at(e)];
            if (atomicCAS(A.res + pt, (uint32_t)Fs, RES_FREE) == (uint32_t)Fs) unreserve_notify(A.atby, A.slotof, A.doom, pt);
          }
        }
        // the slot and its pages go back in the next round's release pass (parallel), not on this block's path
        if (tid == 0) {
          A.ctl[CTL_STEPS] += sl.steps;
          A.ctl[CTL_TX] += 1;
          sl.status = ST_DONE;
        }
        __syncthreads();
        F = Fs + 1;  // the seed is done (its own point stays unmarked, :191)
        done += 1;
        t_slow += gtimer() - ti1;
      } else {
        // ---- fast path: seeds of threads [done, first_special) are dead or tiny ----
        const bool cand = act && tid < first_special && live;
        // cap the batch so that the hash table stays at most half full
        int wsum = cand ? __popc(want) : 0;
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(FULL_MASK, wsum, o);
          if (lane >= o) wsum += v;
        }
        if (lane == 31) sh.warp_sum[tid >> 5] = wsum;
        __syncthreads();
        int incl = wsum;  // inclusive prefix of wants in index order
        for (int w = 0; w < (tid >> 5); ++w) incl += sh.warp_sum[w];
        if (cand && incl > HT / 2) atomicMin(&sh.first_over, tid);
        __syncthreads();
        int seg_hi = sh.first_over < first_special ? sh.first_over : first_special;
        if (seg_hi <= done) seg_hi = done + 1;  // a single seed always fits (K-1 <= 31 wants)
        const bool in_seg = cand && tid < seg_hi;
        // which seeds are about to be marked by a LOWER seed of this batch?
        if (in_seg && want) {
#pragma unroll
          for (int j = 1; j < KM; ++j) {
            if (!((want >> j) & 1u))
              continue;
            const uint32_t id = (uint32_t)ids[j];
            uint32_t h = sweep_hash(id);
            for (;;) {
              const uint32_t old = atomicCAS(hkeys + h, 0xffffffffu, id);
              if (old == 0xffffffffu || old == id) {
                atomicMin(hvals + h, (uint32_t)tid);
                break;
              }
              h = (h + 1) & (HT - 1);
            }
          }
        }
        __syncthreads();
        if (in_seg) {
          uint32_t h = sweep_hash(s);
          for (;;) {
            const uint32_t k = hkeys[h];
            if (k == 0xffffffffu)
              break;
            if (k == s) {
              if (hvals[h] < (uint32_t)tid) atomicMin(&sh.first_conf, tid);
              break;
            }
            h = (h + 1) & (HT - 1);
          }
        }
        __syncthreads();
        const int seg_end = sh.first_conf < seg_hi ? sh.first_conf : seg_hi;  // > done: the first one has nobody below it
        // ---- commit the conflict-free prefix: orphan marks of the tiny transactions (:233 then :238-239) ----
        // Only the owner mark is on the sweeper's path; what the mark means for others (reservation void, holder
        // doomed, alive bit, seed of a waiting grower taken) is logged and applied by a parallel kernel after the
        // sweep -- none of it is needed for correctness (a slot is verified point by point before it commits).
        {
          const bool commit = in_seg && tid < seg_end;
          const int nw = commit ? __popc(want) : 0;
          if (commit) ++ntiny;
          int wincl = nw;
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, wincl, o);
            if (lane >= o) wincl += v;
          }
          const int wtot = __shfl_sync(FULL_MASK, wincl, 31);
          unsigned long long lbase = 0;
          if (wtot > 0) {
            if (lane == 0) lbase = atomicAdd(&S.sc[SC_NLOG], (unsigned long long)wtot);
            lbase = __shfl_sync(FULL_MASK, lbase, 0);
          }
          if (nw) {
            unsigned long long pos = lbase + (unsigned long long)(wincl - nw);
#pragma unroll
            for (int j = 1; j < KM; ++j) {
              if (!((want >> j) & 1u))
                continue;
              const int32_t id = ids[j];
              atomicMin(reinterpret_cast<uint32_t*>(A.state) + id, (uint32_t)i);  // the lower seed owns a shared point
              if (pos < S.marklog_cap) S.marklog[pos] = make_uint2((uint32_t)id, (uint32_t)i);
              ++pos;
            }
          }
        }
        for (int k = tid; k < HT; k += SWEEP_T) {  // (a list of the used slots would need one contended counter)
          hkeys[k] = 0xffffffffu;
          hvals[k] = 0xffffffffu;
        }
        // (no fence: every reader of these marks is in this block, and the barrier orders the block's accesses)
        if (tid == seg_end && valid) sh.last_open = i;  // first seed of the batch that is not finished
        __syncthreads();
        done = seg_end;
        if (done < n_eval) F = sh.last_open;
        t_fast += gtimer() - ti1;
      }
      if (done < n_eval && !stop && state_changed) {
        // ---- refresh: the states of what this thread already knows (one load level) ----
        if (valid && tid >= done) {
          const int32_t st_s = __ldcg(A.state + s);
          int32_t stj[KM];
#pragma unroll
          for (int j = 1; j < KM; ++j) {
            stj[j] = 0;
            if (ids[j] >= 0) stj[j] = __ldcg(A.state + ids[j]);
          }
          slot = __ldcg(A.slotof + i);
          live = st_s == -1;
          want = 0;
          grower = false;
          if (live) {
#pragma unroll
            for (int j = 1; j < KM; ++j)
              if (ids[j] >= 0 && stj[j] == -1) want |= 1u << j;
            grower = __popc(want) == K - 1;
          }
        }
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
    S.sc[SC_T_FRONT] += t_front;
    S.sc[SC_T_SLOW] += t_slow;
    S.sc[SC_T_FAST] += t_fast;
    S.sc[SC_N_SLOW] += n_slow;
    A.ctl[CTL_FRONTIER] = (unsigned long long)F;
    S.sc[SC_SWEEP_ITERS] += iters;
    S.sc[SC_SWEEP_NS] += gtimer() - t_begin;
  }
  // tiny transactions committed: one atomic per warp
  for (int o = 16; o; o >>= 1) ntiny += __shfl_down_sync(FULL_MASK, ntiny, o);
  if (lane == 0 && ntiny) {
    atomicAdd(&S.sc[SC_TINY], ntiny);
    atomicAdd(&A.ctl[CTL_TX], ntiny);
    atomicAdd(&A.ctl[CTL_STEPS], ntiny);
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
  const size_t slot_bytes = (size_t)S.G * sizeof(Slot) + (size_t)S.G * 8 + 256;
  RC_CHECK(dev_ensure(c, c->g_tx, slot_bytes + (size_t)S.G * MAX_PAGES_PER_SLOT * 4 + (size_t)pages * 4 + 64));
  // flag[CMAX+4] | gmask[n] | slotof[n] | atby[n] | alive[words] | doom[n] | hinted[n]
  const int64_t alive_words = (n + 31) / 32 + SWEEP_WORDS + 4;
  RC_CHECK(dev_ensure(c, c->g_spec, (size_t)(CMAX + 4) * 4 + (size_t)n * 14 + (size_t)alive_words * 4 + 256));
  RC_CHECK(dev_ensure(c, c->g_queue, (size_t)pages * PAGE_SIZE * 16 + 256));
  RC_CHECK(dev_ensure(c, c->g_marklog, ((size_t)n + 4096) * sizeof(uint2)));
  S.slots = dptr<Slot>(c->g_tx);
  S.free_ids = reinterpret_cast<uint32_t*>(S.slots + S.G);
  S.ptabs = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(c->g_tx.p) + slot_bytes);
  S.pool.free_pages = S.ptabs + (size_t)S.G * MAX_PAGES_PER_SLOT;
  S.pool.stack_pages = dptr<int2>(c->g_queue);
  S.pool.list_pages = reinterpret_cast<int32_t*>(S.pool.stack_pages + (size_t)pages * PAGE_SIZE);
  S.pool.at_pages = S.pool.list_pages + (size_t)pages * PAGE_SIZE;
  S.flag = dptr<uint32_t>(c->g_spec);
  uint32_t* gmask = S.flag + CMAX + 4;
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
  const size_t sweep_smem = (size_t)HT * 8;
  if (!c->attr_sweep_set) {
    CU_CHECK(c, cudaFuncSetAttribute(spec_sweep_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem));
    CU_CHECK(c, cudaFuncSetAttribute(spec_sweep_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem));
    c->attr_sweep_set = true;
  }

  int64_t F = 0, rounds = 0, stalls = 0, fallbacks = 0;
  const bool dbg = getenv("BSEG_DEBUG") != nullptr;
  cudaEvent_t* pe = c->grow_ev;  // owned by the context: no leak on the error returns below
  float pt[4] = {0, 0, 0, 0};
  for (int k = 0; k < 5; ++k)
    if (!pe[k]) cudaEventCreate(&pe[k]);
  unsigned long long ctl[64] = {0};
  uint32_t* d_ncand = reinterpret_cast<uint32_t*>(&S.sc[SC_NCAND]);
  const unsigned sb = (unsigned)((S.G + GW - 1) / GW);
  // slice length in warp iterations of the head (two-node engine: ~1.6 calls each, a skip batch weighs 3)
  const unsigned long long budget = getenv("BSEG_SLICE") ? strtoull(getenv("BSEG_SLICE"), nullptr, 10) : 2560;
  // the first pass of the sweeper runs before anything is in flight: it stops at the first grower
  while (F < n) {
    const int64_t C = n - F < CMAX ? n - F : CMAX;
    const unsigned gb = (unsigned)ceil_div64(C, TPB);
    S.A.frontier = F;
    cudaEventRecord(pe[0], c->stream);
    if (rounds > 0) {
      spec_mark_release_kernel<<<(S.G + TPB - 1) / TPB, TPB, 0, c->stream>>>(S);
      KLAUNCH_CHECK(c);
      spec_release_entries_kernel<<<dim3(RCH, S.G), TPB, 0, c->stream>>>(S);
      KLAUNCH_CHECK(c);
      spec_release_slots_kernel<<<(S.G + TPB - 1) / TPB, TPB, 0, c->stream>>>(S);
      KLAUNCH_CHECK(c);
      cudaEventRecord(pe[1], c->stream);
      spec_scout_kernel<<<gb, TPB, 0, c->stream>>>(S, F, C);
      KLAUNCH_CHECK(c);
      RC_CHECK(bseg_exclusive_scan_u32(c, S.flag, C, d_ncand));
      spec_assign_kernel<<<gb, TPB, 0, c->stream>>>(S, F, C);
      KLAUNCH_CHECK(c);
      spec_pop_free_kernel<<<1, 1, 0, c->stream>>>(S, F);
      KLAUNCH_CHECK(c);
      cudaEventRecord(pe[2], c->stream);
      spec_grow_kernel<<<sb, GW * 32, 0, c->stream>>>(S, budget);
      KLAUNCH_CHECK(c);
    }
    cudaEventRecord(pe[3], c->stream);
    if (A.K <= 16) spec_sweep_kernel<16><<<1, SWEEP_T, sweep_smem, c->stream>>>(S);
    else spec_sweep_kernel<32><<<1, SWEEP_T, sweep_smem, c->stream>>>(S);
    KLAUNCH_CHECK(c);
    spec_apply_marks_kernel<<<c->num_sms * 4, TPB, 0, c->stream>>>(S);
    KLAUNCH_CHECK(c);
    spec_apply_planes_kernel<<<c->num_sms * 4, TPB, 0, c->stream>>>(S);
    KLAUNCH_CHECK(c);
    spec_reset_log_kernel<<<1, 1, 0, c->stream>>>(S);
    KLAUNCH_CHECK(c);
    cudaEventRecord(pe[4], c->stream);
    ++rounds;
    RC_CHECK(read_back(c, ctl, A.ctl, sizeof(ctl)));
    if (rounds > 1)
      for (int k = 0; k < 4; ++k) {
        float ms = 0;
        cudaEventElapsedTime(&ms, pe[k], pe[k + 1]);
        pt[k] += ms;
      }
    if (ctl[CTL_ERR])
      break;
    const int64_t Fn = (int64_t)ctl[CTL_FRONTIER];
    if (Fn < F)
      return bseg_fail(c, BSEG_E_STATE, "plane grower: frontier moved backwards");
    // no progress and nothing growing at the head for several rounds: the head cannot get a slot (no pages
    // left, or a plane too large for the page table) -- grow it alone in the flat region
    const bool head_has_slot = ctl[8 + SC_HEAD_SLOT] != 0;
    if (Fn == F && rounds > 1 && !head_has_slot) {
      if (++stalls >= 2) {
        launch_grow_seq(c, S.A, true, 1, 1ull << 40, 1);
        KLAUNCH_CHECK(c);
        RC_CHECK(read_back(c, ctl, A.ctl, sizeof(ctl)));
        if (ctl[CTL_ERR])
          break;
        stalls = 0;
        ++fallbacks;
        F = (int64_t)ctl[CTL_FRONTIER];
        continue;
      }
    } else {
      stalls = 0;
    }
    F = Fn;
    if (rounds > 8 * n + 1024)
      return bseg_fail(c, BSEG_E_STATE, "plane grower: no progress");
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
    fprintf(stderr, "[bseg] rounds %lld: release %.1f ms, scout+assign %.1f ms, slices %.1f ms, sweep+apply %.1f ms\n", (long long)rounds,
            pt[0], pt[1], pt[2], pt[3]);
  }
  c->tm.grow_slice_ms = pt[2];
  c->tm.grow_sweep_ms = pt[3];
  if (dbg)
    fprintf(stderr, "[bseg] sweeper: front %.1f ms, slow path %.1f ms (%llu), fast path %.1f ms\n", ctl[8 + SC_T_FRONT] / 1e6,
            ctl[8 + SC_T_SLOW] / 1e6, ctl[8 + SC_N_SLOW], ctl[8 + SC_T_FAST] / 1e6);
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
