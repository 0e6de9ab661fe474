// grow.cuh -- the Broad() step engine shared by the sequential (grow.cu) and speculative
// (grow_spec.cu) plane growers.  Reference semantics: my_function.cpp:180-258 (see grow.cu).
#pragma once
#include "common.cuh"
#include "bseg_arith.h"

static constexpr uint32_t RES_FREE = 0xffffffffu;

struct GrowArgs {
  const int4* pts;
  const double* nrm;
  const int32_t* nbr;
  const uint32_t* inv;
  int32_t* state;
  uint32_t* res;
  int64_t n;
  int K;
  double th_thick, th_dot;
  int64_t th_count;
  int32_t* pool;       // committed lists (CSR), the running list of the sequential engine at its end
  int64_t pool_cap;
  int2* stack;         // DFS frames of the sequential engine
  PlaneRec* planes;
  int64_t planes_cap;
  // control block: [0] next seed (frontier), [1] pool used, [2] planes, [3] steps, [4] error, [5] transactions
  unsigned long long* ctl;
  // speculative engine only (NULL otherwise): per ORIGINAL index
  uint8_t* doom;     // 1 = a lower transaction took a point this one holds, or its seed
  int32_t* slotof;   // grower slot attached to this seed, -1 = none
  // speculative slices: set by the head slot when it finishes, polled by the others
  unsigned long long* stop_flag;  // 1 = stop now; 2 = the head is done: stop once the slice is slice_min_ns old
  unsigned long long slice_min_ns;  // (stop_flag[1] = globaltimer at the start of the slice)
  int64_t frontier;  // first unresolved seed at slice start (everything below is committed); < 0: read ctl[CTL_FRONTIER]
  int flags;         // tuning switches (BSEG_GROW_FLAGS): see GF_*
  const uint8_t* rowdup;  // [n] 1 = the neighbour row names some point twice (dedupe needed, rare)
  uint32_t* atby;    // [n] lowest in-flight transaction that assumed the point taken (early notification)
};
enum { GF_ROW_L1 = 1, GF_ROW_L2 = 2, GF_EARLY_POP = 4, GF_STATE_NC = 8, GF_ROWDUP = 16, GF_FASTDIV = 32, GF_NOPAIR = 64, GF_NOSKIP = 128, GF_SKIP_EAGER = 256, GF_EXACT_MODEL = 512, GF_TIMING = 1024, GF_BG = 2048 /* background slice: see tx_run_pair */, GF_NEVER = 1 << 30 /* never set */ };

// ---- "assumed taken" list of a speculative transaction --------------------------------------------------
// A point that is free in the committed state and passes the geometric tests, but is reserved by a LOWER
// in-flight transaction (a grower, or a tiny transaction that published its orphan marks ahead of the
// sweeper), is treated as taken and appended to the slot's AT list.  When the sweeper reaches the
// transaction every lower one has been committed, so the result is the sequential one iff every AT point
// is taken by then -- that is what the sweeper verifies before committing (grow_spec.cu).
// Reservations below the slice's frontier belong to transactions the sweeper has passed: they are stale
// (a dead tiny transaction never clears its hints) and count as free.
// Early notification: atby[pt] names the lowest transaction that assumed pt taken; a reservation that is
// dropped WITHOUT the point being taken (retracted hint, released or rolled-back grower) dooms it right
// away, so it re-runs ahead of the sweeper instead of failing the verification at the head.
__device__ __forceinline__ void unreserve_notify(uint32_t* atby, const int32_t* slotof, uint8_t* doom, int64_t p)
{
  if (__ldcg(atby + p) == 0xffffffffu)
    return;
  const uint32_t a = atomicExch(atby + p, 0xffffffffu);
  if (a != 0xffffffffu && __ldcg(slotof + a) >= 0) doom[a] = 1;
}

// owner and reservation of point p in one 8-byte load (.x = state, .y = reservation): the two live in one sector
__device__ __forceinline__ int2 ld_state_res(const GrowArgs& A, int64_t p)
{
  return __ldcg(reinterpret_cast<const int2*>(A.state) + p);
}

enum { CTL_FRONTIER = 0, CTL_POOL = 1, CTL_PLANES = 2, CTL_STEPS = 3, CTL_ERR = 4, CTL_TX = 5 };

// model of the running plane, replicated on every lane of the warp
struct Model {
  double mn0, mn1, mn2;  // cur_normal
  int32_t mc0, mc1, mc2; // cur_center
  double sn0, sn1, sn2;  // running sum of normals
  uint32_t sc0, sc1, sc2;// running (wrapping) sum of positions
};

__device__ __forceinline__ void model_init(Model& m, const int4& p, const double* __restrict__ nr)
{
  m.mn0 = nr[0]; m.mn1 = nr[1]; m.mn2 = nr[2];
  m.mc0 = p.x; m.mc1 = p.y; m.mc2 = p.z;
  m.sn0 = 0.0 + m.mn0; m.sn1 = 0.0 + m.mn1; m.sn2 = 0.0 + m.mn2;
  m.sc0 = (uint32_t)p.x; m.sc1 = (uint32_t)p.y; m.sc2 = (uint32_t)p.z;
}

// my_function.cpp:227-230 for one neighbour
__device__ __forceinline__ bool geo_test(const Model& m, const int4& p, double n0, double n1, double n2,
                                         double th_thick, double th_dot)
{
  int32_t p0 = (int32_t)((uint32_t)p.x - (uint32_t)m.mc0);
  int32_t p1 = (int32_t)((uint32_t)p.y - (uint32_t)m.mc1);
  int32_t p2 = (int32_t)((uint32_t)p.z - (uint32_t)m.mc2);
  double dist = bseg_fabs((double)p0 * m.mn0 + (double)p1 * m.mn1 + (double)p2 * m.mn2);
  double dot = m.mn0 * n0 + m.mn1 * n1 + m.mn2 * n2;
  return dist <= th_thick && dot >= th_dot;
}

// cur_center /= size (my_function.cpp:250): the int32 sum is converted to size_t first (PCCMath.h:227-235)
__device__ __forceinline__ int32_t center_div(uint32_t sc, uint32_t len)
{
  if ((int32_t)sc >= 0)
    return (int32_t)(sc / len);  // same quotient as the 64-bit form, 32-bit divide
  return (int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)sc) / (uint64_t)len);
}

// my_function.cpp:241-250 after the accepted points were added to the running sums
__device__ __forceinline__ void model_update(Model& m, int64_t len)
{
  double nn = bseg_sqrt((m.sn0 * m.sn0) + (m.sn1 * m.sn1) + (m.sn2 * m.sn2));
  m.mn0 = m.sn0 / nn; m.mn1 = m.sn1 / nn; m.mn2 = m.sn2 / nn;
  m.mc0 = center_div(m.sc0, (uint32_t)len);
  m.mc1 = center_div(m.sc1, (uint32_t)len);
  m.mc2 = center_div(m.sc2, (uint32_t)len);
}

// The same update with the three quotients sharing one reciprocal -- bit-identical results:
//   fp64: Markstein's theorem -- r = RN(1/b), q0 = RN(a*r), rem = a - b*q0 (exact, FMA), RN(q0 + rem*r) is the
//         correctly rounded a/b unless b's mantissa is all ones; operands outside a safe exponent window
//         (zero, subnormal, inf, NaN: Q9) take the true division;
//   u32:  floor(a/b) from trunc(a * RN(1/b)) is off by at most one; the remainder tells which way.
__device__ __forceinline__ bool div_safe_operand(double x)
{
  const int e = (int)((bseg_d2u(x) >> 52) & 0x7ffu);
  return e > 523 && e < 1523;
}
__device__ __forceinline__ void model_update_fast(Model& m, int64_t len)
{
  const double nn = bseg_sqrt((m.sn0 * m.sn0) + (m.sn1 * m.sn1) + (m.sn2 * m.sn2));
  const bool all_ones = (bseg_d2u(nn) & 0xfffffffffffffULL) == 0xfffffffffffffULL;
  if (div_safe_operand(nn) && !all_ones && div_safe_operand(m.sn0) && div_safe_operand(m.sn1) &&
      div_safe_operand(m.sn2)) {
    const double r = __drcp_rn(nn);
    const double q0 = __dmul_rn(m.sn0, r), q1 = __dmul_rn(m.sn1, r), q2 = __dmul_rn(m.sn2, r);
    m.mn0 = __fma_rn(__fma_rn(-q0, nn, m.sn0), r, q0);
    m.mn1 = __fma_rn(__fma_rn(-q1, nn, m.sn1), r, q1);
    m.mn2 = __fma_rn(__fma_rn(-q2, nn, m.sn2), r, q2);
  } else {
    m.mn0 = m.sn0 / nn; m.mn1 = m.sn1 / nn; m.mn2 = m.sn2 / nn;
  }
  const uint32_t l = (uint32_t)len;
  if ((int32_t)(m.sc0 | m.sc1 | m.sc2) >= 0) {  // all three sums below 2^31 (always, before Q6 strikes)
    const double rl = __drcp_rn((double)l);
    uint32_t q, rem;
    q = (uint32_t)__dmul_rn((double)m.sc0, rl); rem = m.sc0 - q * l;
    m.mc0 = (int32_t)((int32_t)rem < 0 ? q - 1 : (rem >= l ? q + 1 : q));
    q = (uint32_t)__dmul_rn((double)m.sc1, rl); rem = m.sc1 - q * l;
    m.mc1 = (int32_t)((int32_t)rem < 0 ? q - 1 : (rem >= l ? q + 1 : q));
    q = (uint32_t)__dmul_rn((double)m.sc2, rl); rem = m.sc2 - q * l;
    m.mc2 = (int32_t)((int32_t)rem < 0 ? q - 1 : (rem >= l ? q + 1 : q));
  } else {
    m.mc0 = center_div(m.sc0, l);
    m.mc1 = center_div(m.sc1, l);
    m.mc2 = center_div(m.sc2, l);
  }
}

// ---- the model between two decisions (speculative two-node engine) ------------------------------------------
// The reference renormalises the model after every accepting call: a correctly rounded sqrt and three correctly
// rounded divisions, ~1500 cycles of dependent fp64 work on the grower's critical path -- only to compare two
// numbers with thresholds that almost no point comes near.  The engine therefore carries cur_normal as
// sn * rsqrt(sn.sn), which is within 2^-50 (relative) of the exact quotients, and decides with it whenever both
// test values are clear of their thresholds by far more than that can move them (|dot - th_dot| > 1e-9; the
// distance term sums three products of magnitude < 2^24, so its error is < 1e-7: margin 1e-4).  A test value
// inside a margin makes the warp compute the exact model from the (always exact) running sums and decide again;
// the exact model is also what a finished transaction reports.  cur_center (an integer quotient) is always exact.
__device__ __forceinline__ bool geo_test_margin(const Model& m, const int4& p, double n0, double n1, double n2,
                                                double th_thick, double th_dot, bool& near)
{
  int32_t p0 = (int32_t)((uint32_t)p.x - (uint32_t)m.mc0);
  int32_t p1 = (int32_t)((uint32_t)p.y - (uint32_t)m.mc1);
  int32_t p2 = (int32_t)((uint32_t)p.z - (uint32_t)m.mc2);
  double dist = bseg_fabs((double)p0 * m.mn0 + (double)p1 * m.mn1 + (double)p2 * m.mn2);
  double dot = m.mn0 * n0 + m.mn1 * n1 + m.mn2 * n2;
  near = bseg_fabs(dist - th_thick) <= 1e-4 || bseg_fabs(dot - th_dot) <= 1e-9;
  return dist <= th_thick && dot >= th_dot;
}

// cur_center /= size with a float-seeded reciprocal (one Newton step: relative error < 2^-43, the quotient is
// off by at most one and the remainder tells which way) -- same result as center_div
__device__ __forceinline__ void center_update(Model& m, uint32_t l)
{
  if ((int32_t)(m.sc0 | m.sc1 | m.sc2) >= 0) {  // all three sums below 2^31 (always, before Q6 strikes)
    const double dl = (double)l;
    const double r0 = (double)__frcp_rn((float)l);
    const double rl = __fma_rn(r0, __fma_rn(-dl, r0, 1.0), r0);
    uint32_t q, rem;
    q = (uint32_t)__dmul_rn((double)m.sc0, rl); rem = m.sc0 - q * l;
    m.mc0 = (int32_t)((int32_t)rem < 0 ? q - 1 : (rem >= l ? q + 1 : q));
    q = (uint32_t)__dmul_rn((double)m.sc1, rl); rem = m.sc1 - q * l;
    m.mc1 = (int32_t)((int32_t)rem < 0 ? q - 1 : (rem >= l ? q + 1 : q));
    q = (uint32_t)__dmul_rn((double)m.sc2, rl); rem = m.sc2 - q * l;
    m.mc2 = (int32_t)((int32_t)rem < 0 ? q - 1 : (rem >= l ? q + 1 : q));
  } else {
    m.mc0 = center_div(m.sc0, l);
    m.mc1 = center_div(m.sc1, l);
    m.mc2 = center_div(m.sc2, l);
  }
}

// returns false when the sums are outside the range the error bound was derived for (the caller goes exact)
__device__ __forceinline__ bool model_update_approx(Model& m, int64_t len)
{
  const double ss = (m.sn0 * m.sn0) + (m.sn1 * m.sn1) + (m.sn2 * m.sn2);
  if (!div_safe_operand(ss) || !div_safe_operand(m.sn0) || !div_safe_operand(m.sn1) || !div_safe_operand(m.sn2))
    return false;
  const double r = rsqrt(ss);
  m.mn0 = m.sn0 * r; m.mn1 = m.sn1 * r; m.mn2 = m.sn2 * r;
  center_update(m, (uint32_t)len);
  return true;
}

// add the accepted lanes' normals / positions to the running sums in neighbour order
__device__ __forceinline__ void model_accumulate(Model& m, uint32_t acc, const int4& p, double n0, double n1, double n2)
{
  while (acc) {  // two accepted lanes per trip: their shuffles are in flight together, the sums stay in order
    const int b = __ffs(acc) - 1;
    acc &= acc - 1;
    const bool two = acc != 0;
    const int c = two ? __ffs(acc) - 1 : b;
    acc &= acc - 1;
    const double x0 = __shfl_sync(FULL_MASK, n0, b), x1 = __shfl_sync(FULL_MASK, n1, b), x2 = __shfl_sync(FULL_MASK, n2, b);
    const uint32_t u0 = (uint32_t)__shfl_sync(FULL_MASK, p.x, b), u1 = (uint32_t)__shfl_sync(FULL_MASK, p.y, b),
                   u2 = (uint32_t)__shfl_sync(FULL_MASK, p.z, b);
    const double y0 = __shfl_sync(FULL_MASK, n0, c), y1 = __shfl_sync(FULL_MASK, n1, c), y2 = __shfl_sync(FULL_MASK, n2, c);
    const uint32_t v0 = (uint32_t)__shfl_sync(FULL_MASK, p.x, c), v1 = (uint32_t)__shfl_sync(FULL_MASK, p.y, c),
                   v2 = (uint32_t)__shfl_sync(FULL_MASK, p.z, c);
    m.sn0 += x0; m.sn1 += x1; m.sn2 += x2;
    m.sc0 += u0; m.sc1 += u1; m.sc2 += u2;
    if (two) {
      m.sn0 += y0; m.sn1 += y1; m.sn2 += y2;
      m.sc0 += v0; m.sc1 += v1; m.sc2 += v2;
    }
  }
}

// keep only the lowest lane of every group of accepted lanes that name the same point
// (a row with repeated ids: the second occurrence sees planeIdx already set, :226)
__device__ __forceinline__ bool dedupe(bool ok, int32_t id)
{
  uint32_t same = __match_any_sync(FULL_MASK, ok ? id : -1 - (int)(threadIdx.x & 31));
  return ok && ((same & lanemask_lt()) == 0);
}

__device__ __forceinline__ void prefetch_l2(const void* p)
{
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l1(const void* p)
{
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// speculative slices: the head finishing ends the slice for everybody -- but not before the slice is slice_min_ns old
// (a round has a fixed cost: release, scout, sweep; several short planes in a row share one)
__device__ __forceinline__ bool slice_over(const GrowArgs& A)
{
  const unsigned long long st = *(volatile unsigned long long*)A.stop_flag;
  if (st == 0)
    return false;
  if (st == 1)
    return true;
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (now - ((volatile unsigned long long*)A.stop_flag)[1] < A.slice_min_ns)
    return false;
  *(volatile unsigned long long*)A.stop_flag = 1ull;
  return true;
}

// ---- list / frame storage ---------------------------------------------------------------------------
// flat: the sequential engine's region at the end of the committed pool
struct FlatStore {
  int32_t* list;
  int2* stack;
  __device__ __forceinline__ bool reserve(int64_t, int) { return true; }
  __device__ __forceinline__ void put(int64_t e, int32_t id) { list[e] = id; }
  __device__ __forceinline__ int32_t get(int64_t e) const { return list[e]; }
  __device__ __forceinline__ void push(int64_t sp, int2 f) { stack[sp] = f; }
  __device__ __forceinline__ int2 pop(int64_t sp) const { return stack[sp]; }
  __device__ __forceinline__ void put_at(int64_t, int32_t) {}
};

// paged: grower slots of the speculative engine share one page pool
constexpr int PAGE_SHIFT = 11;
constexpr int PAGE_SIZE = 1 << PAGE_SHIFT;  // entries per page
constexpr int MAX_PAGES_PER_SLOT = 2048;    // 4 M entries per plane; larger planes run in the flat region

struct PagePool {
  int32_t* list_pages;   // [n_pages][PAGE_SIZE]
  int2* stack_pages;     // [n_pages][PAGE_SIZE]
  int32_t* at_pages;     // [n_pages][PAGE_SIZE] assumed-taken points
  uint32_t* free_pages;  // stack of free page ids
  unsigned long long* n_free;
};

struct PagedStore {
  PagePool pool;
  uint32_t* ptab;       // this slot's page table [MAX_PAGES_PER_SLOT]
  int32_t* n_pages;     // pages currently owned by the slot
  int have_cached;      // this warp is the only writer while it runs: a register copy saves a load per step
  // make entries [0, upto) addressable; one lane allocates, the warp learns the result
  __device__ __forceinline__ bool reserve(int64_t upto, int lane)
  {
    int have = have_cached;
    const int need = (int)((upto + PAGE_SIZE - 1) >> PAGE_SHIFT);
    if (need <= have)
      return true;
    int ok = 1;
    if (lane == 0) {
      while (have < need) {
        if (have >= MAX_PAGES_PER_SLOT) { ok = 0; break; }
        unsigned long long nf = atomicAdd(pool.n_free, ~0ull);  // --n_free, returns old
        if ((long long)nf <= 0) {
          atomicAdd(pool.n_free, 1ull);
          ok = 0;
          break;
        }
        ptab[have++] = pool.free_pages[nf - 1];
      }
      *n_pages = have;
    }
    ok = __shfl_sync(FULL_MASK, ok, 0);
    have_cached = __shfl_sync(FULL_MASK, have, 0);
    __syncwarp();
    return ok != 0;
  }
  __device__ __forceinline__ size_t at(int64_t e) const
  {
    return ((size_t)ptab[e >> PAGE_SHIFT] << PAGE_SHIFT) + (size_t)(e & (PAGE_SIZE - 1));
  }
  __device__ __forceinline__ void put(int64_t e, int32_t id) { pool.list_pages[at(e)] = id; }
  __device__ __forceinline__ int32_t get(int64_t e) const { return pool.list_pages[at(e)]; }
  __device__ __forceinline__ void push(int64_t sp, int2 f) { pool.stack_pages[at(sp)] = f; }
  __device__ __forceinline__ int2 pop(int64_t sp) const { return pool.stack_pages[at(sp)]; }
  __device__ __forceinline__ void put_at(int64_t k, int32_t id) { pool.at_pages[at(k)] = id; }
  __device__ __forceinline__ int32_t get_at(int64_t k) const { return pool.at_pages[at(k)]; }
};

// ---- one transaction's DFS state -------------------------------------------------------------------
struct TxState {
  Model m;
  int64_t len, sp, top_cur, top_end;
  int64_t n_at;  // entries of the assumed-taken list (speculative engine)
  uint32_t node;
  int have_top, depth0;
  int model_exact;  // m.mn* are the reference's correctly rounded quotients (else within ~2^-50 of them, see model_update_approx)
};

enum TxOutcome {
  TX_RUNNING = 0,      // budget exhausted, resumable
  TX_FINISHED = 1,     // DFS complete (commit or roll back by size)
  TX_FAILED = 2,       // depth-0 failure: orphan marks stay, no plane
  TX_DOOMED = 3,       // overlapped a lower in-flight transaction
  TX_OVERFLOW = 4,     // storage exhausted
  TX_IS_GROWER = 5     // sequential engine asked to leave growers alone: nothing was marked
};

enum { MODE_SEQ = 0, MODE_SEQ_NOTIFY = 1, MODE_SPEC = 2 };

__device__ __forceinline__ void tx_begin(TxState& t, const GrowArgs& A, uint32_t seed_s)
{
  model_init(t.m, __ldg(A.pts + seed_s), A.nrm + 3 * (int64_t)seed_s);
  t.len = 1;
  t.n_at = 0;
  t.sp = 0;
  t.top_cur = 0;
  t.top_end = 0;
  t.node = seed_s;
  t.have_top = 0;
  t.depth0 = 1;
  t.model_exact = 1;
}

// Runs Broad() calls of transaction `seed_i` until it ends or `budget` calls were made.
//   MODE_SEQ         committed state only; accepted points are marked in `state` at once
//   MODE_SEQ_NOTIFY  same, and in-flight speculative transactions touching them are doomed
//   MODE_SPEC        marks are reservations (atomicMin on res); a lower reservation = "assume taken"
// Latency chain of one step (the grower is latency-bound): row of the next node (L1: every neighbour's
// row is prefetched when the neighbour is gathered) -> gather of its neighbours' state / reservation /
// position / normal (L2) -> tests.  The reservation atomics are NOT on the chain in MODE_SPEC: the
// decision uses the reservation value that was gathered, the atomic's result is inspected one step later
// and only matters when a lower transaction slipped in between (then this one gives up and is re-run).
template <int MODE, int KT, class Store>
__device__ TxOutcome tx_run(const GrowArgs& A, Store& st, TxState& t, int64_t seed_i, unsigned long long budget,
                            bool leave_growers, int lane, unsigned long long& steps_out)
{
  const int K = KT ? KT : A.K;  // KT: compile-time row length (15 = the reference's), 0 = run time
  const uint32_t me = (uint32_t)seed_i;
  const bool row_l1 = (A.flags & GF_ROW_L1) != 0, row_l2 = (A.flags & GF_ROW_L2) != 0;
  const bool early_pop = (A.flags & GF_EARLY_POP) != 0, state_nc = (A.flags & GF_STATE_NC) != 0;
  const bool use_rowdup = (A.flags & GF_ROWDUP) != 0, fastdiv = (A.flags & GF_FASTDIV) != 0;
  const uint32_t fr = MODE == MODE_SPEC ? (A.frontier < 0 ? (uint32_t)__ldcg(A.ctl + CTL_FRONTIER) : (uint32_t)A.frontier) : 0u;
  bool has_dup = true;  // of the row being tested
  if (use_rowdup) has_dup = __ldg(A.rowdup + t.node) != 0;
  unsigned long long steps = 0;
  TxOutcome out = TX_RUNNING;
  // neighbour `lane` of the node: id, and its state / reservation / position / normal
  int32_t id = -1, stt = 0;
  uint32_t rs = RES_FREE;
  bool mine = false;  // accepted by this transaction in the previous step (its reservation may still be in flight)
  int4 p = make_int4(0, 0, 0, 0);
  double n0 = 0, n1 = 0, n2 = 0;
  if (lane >= 1 && lane < K)
    id = __ldg(A.nbr + (int64_t)t.node * K + lane);
  if (id >= 0) {
    // inside a speculative slice nobody writes `state` (the sweeper is a different kernel): L1 may keep it
    stt = (MODE == MODE_SPEC && state_nc) ? __ldg(A.state + 2 * (int64_t)(id)) : __ldcg(A.state + 2 * (int64_t)(id));
    if (MODE == MODE_SPEC) rs = __ldcg(A.res + 2 * (int64_t)(id));
    p = __ldg(A.pts + id);
    const double* nr = A.nrm + 3 * (int64_t)id;
    n0 = __ldg(nr); n1 = __ldg(nr + 1); n2 = __ldg(nr + 2);
    if (row_l1) {
      prefetch_l1(A.nbr + (int64_t)id * K);
      prefetch_l1(A.nbr + (int64_t)id * K + (K - 1));
    }
  }
  while (steps < budget) {
    if (!st.reserve((t.len > t.n_at ? t.len : t.n_at) + 2 * K, lane)) {  // before anything is marked
      out = TX_OVERFLOW;
      break;
    }
    ++steps;
    // the node visited if nothing is accepted now: fetch it early, it is two dependent loads away
    int32_t pop_id = -1;
    const int64_t pop_at = t.top_cur;
    if (early_pop && t.have_top && t.top_cur < t.top_end) pop_id = st.get(t.top_cur);
    bool ok = id >= 0 && stt == -1 && (MODE != MODE_SPEC || (rs != me && !mine)) &&
              geo_test(t.m, p, n0, n1, n2, A.th_thick, A.th_dot);
    if (has_dup) ok = dedupe(ok, id);
    if (MODE == MODE_SPEC && (A.flags & GF_BG)) {  // background slice: nothing is taken from anybody (see tx_run_pair)
      const bool contested = ok && rs != RES_FREE && rs > me;
      if (__any_sync(FULL_MASK, contested)) {
        --steps;
        break;
      }
    }
    if (leave_growers && t.depth0) {
      if (__popc(__ballot_sync(FULL_MASK, ok)) == K - 1) {
        out = TX_IS_GROWER;
        --steps;
        break;
      }
    }
    bool relied = false;  // a lower in-flight transaction holds the point: assume taken
    uint32_t old = RES_FREE;
    bool fired = false;
    if (ok) {
      if (MODE == MODE_SPEC) {
        if (rs < fr) {  // stale reservation of a transaction the sweeper has passed: free; replace it (rare)
          uint32_t cur = rs;
          for (;;) {
            if (cur < fr) {
              const uint32_t seen = atomicCAS(A.res + 2 * (int64_t)(id), cur, me);
              if (seen == cur) {
                old = RES_FREE;
                break;
              }
              cur = seen;
            } else {
              old = atomicMin(A.res + 2 * (int64_t)(id), me);
              if (old >= fr)
                break;
              cur = old;
            }
          }
          if (old < me) {
            ok = false;
            relied = true;
          } else {
            if (old != RES_FREE) A.doom[old] = 1;
            if (p.w > (int32_t)me) A.doom[p.w] = 1;
          }
        } else if (rs < me) {
          ok = false;
          relied = true;
        } else {
          old = atomicMin(A.res + 2 * (int64_t)(id), me);  // result looked at after the next gather is on its way
          fired = true;
        }
      } else {
        A.state[2 * (int64_t)id] = (int32_t)seed_i;
        if (MODE == MODE_SEQ_NOTIFY) {
          const uint32_t o = atomicMin(A.res + 2 * (int64_t)(id), me);
          if (o != RES_FREE && o > me) A.doom[o] = 1;
          if (p.w > (int32_t)me) A.doom[p.w] = 1;
        }
      }
    }
    if (MODE == MODE_SPEC) {
      const uint32_t rel = __ballot_sync(FULL_MASK, relied);
      if (rel) {
        if (relied) {
          st.put_at(t.n_at + __popc(rel & lanemask_lt()), id);
          atomicMin(A.atby + id, me);  // whoever un-reserves the point without taking it tells us at once
        }
        t.n_at += __popc(rel);
      }
    }
    const uint32_t acc = __ballot_sync(FULL_MASK, ok);
    const int cnt = __popc(acc);
    if (ok)
      st.put(t.len + __popc(acc & lanemask_lt()), id);
    __syncwarp();
    if (t.depth0 && cnt < K - 1) {
      t.len += cnt;
      out = TX_FAILED;  // :238-239
      break;
    }
    t.depth0 = 0;
    // ---- DFS bookkeeping: which node is next (:252-255) ----
    const int64_t s0 = t.len;
    t.len += cnt;
    uint32_t next = 0;
    bool have_next = false;
    if (cnt > 0) {
      if (t.have_top && t.top_cur < t.top_end) {
        if (lane == 0) st.push(t.sp, make_int2((int)t.top_cur, (int)t.top_end));
        ++t.sp;
      }
      t.top_cur = s0 + 1;  // the first accepted point is visited right away
      t.top_end = t.len;
      t.have_top = 1;
      next = (uint32_t)__shfl_sync(FULL_MASK, id, __ffs(acc) - 1);
      have_next = true;
      // every accepted point gets its own Broad() later: pull its row towards L2 now
      if (row_l2 && ok) prefetch_l2(A.nbr + (int64_t)id * K);
    } else {
      while (t.have_top && t.top_cur == t.top_end) {
        if (t.sp > 0) {
          --t.sp;
          __syncwarp();
          const int2 f = st.pop(t.sp);
          t.top_cur = f.x;
          t.top_end = f.y;
        } else {
          t.have_top = 0;
        }
      }
      if (t.have_top) {
        next = (uint32_t)((pop_id >= 0 && t.top_cur == pop_at) ? pop_id : st.get(t.top_cur));
        ++t.top_cur;
        have_next = true;
      }
    }
    // ---- software pipeline: next row, running sums, next gather, model update ----
    int32_t nid = -1;
    if (have_next && lane >= 1 && lane < K)
      nid = __ldg(A.nbr + (int64_t)next * K + lane);
    bool ndup = true;
    if (use_rowdup && have_next) ndup = __ldg(A.rowdup + next) != 0;
    model_accumulate(t.m, acc, p, n0, n1, n2);
    int32_t nstt = 0;
    uint32_t nrs = RES_FREE;
    bool nmine = false;
    int4 np = make_int4(0, 0, 0, 0);
    double m0 = 0, m1 = 0, m2 = 0;
    if (nid >= 0) {
      nstt = (MODE == MODE_SPEC && state_nc) ? __ldg(A.state + 2 * (int64_t)(nid)) : __ldcg(A.state + 2 * (int64_t)(nid));
      if (MODE == MODE_SPEC) nrs = __ldcg(A.res + 2 * (int64_t)(nid));
      np = __ldg(A.pts + nid);
      const double* nr = A.nrm + 3 * (int64_t)nid;
      m0 = __ldg(nr); m1 = __ldg(nr + 1); m2 = __ldg(nr + 2);
      if (row_l1) {
        prefetch_l1(A.nbr + (int64_t)nid * K);
        prefetch_l1(A.nbr + (int64_t)nid * K + (K - 1));
      }
    }
    if (MODE == MODE_SPEC) {  // points accepted a moment ago are ours whatever the gathered reservation says
      uint32_t a = acc;
      while (a) {
        const int b = __ffs(a) - 1;
        a &= a - 1;
        nmine |= nid == __shfl_sync(FULL_MASK, id, b);
      }
    }
    if (fastdiv) model_update_fast(t.m, t.len);
    else model_update(t.m, t.len);
    if (MODE == MODE_SPEC) {
      bool race = false;
      if (fired) {
        if (old < me && old >= fr) race = true;         // a lower transaction reserved it in between
        else if (old != RES_FREE && old >= fr) A.doom[old] = 1;  // stolen from a higher transaction: it is void
        if (p.w > (int32_t)me) A.doom[p.w] = 1;         // the transaction seeded at this point lost its seed
      }
      if (__any_sync(FULL_MASK, race)) {
        out = TX_DOOMED;
        break;
      }
    }
    if (!have_next) {
      out = TX_FINISHED;
      break;
    }
    t.node = next;
    id = nid; stt = nstt; rs = nrs; mine = nmine; p = np; n0 = m0; n1 = m1; n2 = m2;
    has_dup = ndup;
    if (MODE == MODE_SPEC && (steps & 7) == 0) {
      if (((volatile uint8_t*)A.doom)[seed_i]) {
        out = TX_DOOMED;
        break;
      }
      if (slice_over(A)) break;  // the head finished: let the sweeper commit
    }
  }
  steps_out += steps;
  return out;
}

// ---- two nodes per warp step (speculative engine, K <= 16) ------------------------------------------------
// Three Broad() calls out of four accept nothing (the neighbours are already in the plane): the model is
// unchanged by such a call (same sums, same length), and the node visited next is simply the next entry of
// the current DFS frame.  So half-warp 0 tests the node A that is due, half-warp 1 the node B that follows
// if A accepts nothing, both against the same model.  When A accepts nothing B's outcome IS the next call
// of the reference, otherwise B's tests are discarded (B stays in its frame).  Exactly the calls of the
// sequential run, in the same order; only their latency is shared.
__device__ __forceinline__ void tx_reserve_lane(const GrowArgs& A, int32_t id, uint32_t rs, uint32_t me, uint32_t fr,
                                                int32_t pw, bool& ok, bool& relied, bool& fired, uint32_t& old)
{
  // lane wants `id`: decide from the gathered reservation `rs` (see tx_run)
  if (rs < fr) {  // stale reservation of a transaction the sweeper has passed: free; replace it (rare)
    uint32_t cur = rs;
    for (;;) {
      if (cur < fr) {
        const uint32_t seen = atomicCAS(A.res + 2 * (int64_t)(id), cur, me);
        if (seen == cur) {
          old = RES_FREE;
          break;
        }
        cur = seen;
      } else {
        old = atomicMin(A.res + 2 * (int64_t)(id), me);
        if (old >= fr)
          break;
        cur = old;
      }
    }
    if (old < me) {
      ok = false;
      relied = true;
    } else {
      if (old != RES_FREE) A.doom[old] = 1;
      if (pw > (int32_t)me) A.doom[pw] = 1;
    }
    old = RES_FREE;
  } else if (rs < me) {
    ok = false;
    relied = true;
  } else {
    old = atomicMin(A.res + 2 * (int64_t)(id), me);  // result looked at after the next gather is on its way
    fired = true;
  }
}

// ---- skipping runs of calls that accept nothing (speculative engine) ----------------------------------------
// Three Broad() calls out of four accept nothing, and they come in RUNS (the DFS unwinding over a finished
// region: 88 % of them sit in runs of 32 and more, the last run of a large plane is half of its calls).  Such a
// call changes nothing: no mark, no list entry, the model keeps its sums.  The nodes the DFS visits while
// nothing is accepted are known in advance: the rest of the current frame, then the frames below it on the
// stack (my_function.cpp:252-255).  So one warp step looks at the next 32 of them at once, one node per lane,
// all against the current model; the calls before the first node that WANTS a point (free in the committed
// state, not ours, passes the geometric tests -- whatever its reservation says) are exactly the calls the
// reference makes next, they are counted and the frames advanced past them.  The wanting node is left to the
// regular step.  Conservative by construction: stale reservation data can only end a run early.
struct SkipScratch {
  int32_t id[32 * 16];    // (node, neighbour) pairs that need the geometric tests: neighbour id ...
  uint8_t owner[32 * 16]; // ... and the lane holding the node
  uint32_t hit;           // lanes whose node wants a point
};

__device__ __forceinline__ uint32_t ld_cg_u32(const void* p)
{
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// lanes whose node (one per lane, `act` lanes only) wants a point
template <int KT>
__device__ __forceinline__ uint32_t skip_eval(const GrowArgs& A, const Model& m, uint32_t me, bool act, uint32_t node,
                                              int lane, SkipScratch* ss, unsigned long long* pairs)
{
  const int K = KT ? KT : A.K;
  // ---- phase 1: the node's row, then state and reservation of ALL its neighbours in one round trip ----
  // (every load is issued before the first test: a padded column reads the node's own entry)
  const uint32_t nd = act ? node : 0u;
  const int32_t* row = A.nbr + (int64_t)nd * K;
  int32_t ids[16];
#pragma unroll
  for (int k = 1; k < 16; ++k) ids[k] = (k < K) ? __ldg(row + k) : -1;
  uint32_t stt[16], rs[16];
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    const uint32_t sid = ids[k] >= 0 ? (uint32_t)ids[k] : nd;
    stt[k] = ld_cg_u32(A.state + 2 * (int64_t)(sid));  // (two 32-bit loads of one sector: a 64-bit one costs registers here)
    rs[k] = ld_cg_u32(A.res + 2 * (int64_t)(sid));
  }
  // Every test below is made to depend on ALL the loads above (a mask that is zero at run time but unknown to
  // the compiler): ptxas would otherwise sink each load next to its test to save registers and serialise the
  // 28 round trips.
  uint32_t mix = 0;
#pragma unroll
  for (int k = 1; k < 16; ++k) mix ^= stt[k] ^ rs[k];
  const uint32_t z = mix & (uint32_t)(A.flags & GF_NEVER);
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    stt[k] ^= z;
    rs[k] ^= z;
  }
  uint32_t cand = 0;
#pragma unroll
  for (int k = 1; k < 16; ++k)
    if (ids[k] >= 0 && stt[k] == 0xffffffffu && rs[k] != me) cand |= 1u << k;
  if (!act) cand = 0;
  // ---- phase 2: geometric tests of the free neighbours, spread over the lanes ----
  const int nc = __popc(cand);
  int cincl = nc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(FULL_MASK, cincl, o);
    if (lane >= o) cincl += v;
  }
  const int C = __shfl_sync(FULL_MASK, cincl, 31);
  uint32_t hit = 0;
  if (C) {
    if (lane == 0) ss->hit = 0;
    int w = cincl - nc;
#pragma unroll
    for (int k = 1; k < 16; ++k)
      if ((cand >> k) & 1u) {
        ss->id[w] = ids[k];
        ss->owner[w] = (uint8_t)lane;
        ++w;
      }
    __syncwarp();
    for (int r0 = 0; r0 < C; r0 += 64) {  // two pairs per lane and trip
      const int ra = r0 + lane, rb = r0 + 32 + lane;
      const bool va = ra < C, vb = rb < C;
      const int32_t ia = ss->id[va ? ra : 0], ib = ss->id[vb ? rb : 0];
      const int4 pa = __ldg(A.pts + ia), pb = __ldg(A.pts + ib);
      const double* na = A.nrm + 3 * (int64_t)ia;
      const double* nb = A.nrm + 3 * (int64_t)ib;
      const double a0 = __ldg(na), a1 = __ldg(na + 1), a2 = __ldg(na + 2);
      const double b0 = __ldg(nb), b1 = __ldg(nb + 1), b2 = __ldg(nb + 2);
      bool near_a = false, near_b = false;  // inside a margin counts as a hit: the regular step decides exactly
      const bool ga = geo_test_margin(m, pa, a0, a1, a2, A.th_thick, A.th_dot, near_a);
      const bool gb = geo_test_margin(m, pb, b0, b1, b2, A.th_thick, A.th_dot, near_b);
      if (va && (ga || near_a)) atomicOr(&ss->hit, 1u << ss->owner[ra]);
      if (vb && (gb || near_b)) atomicOr(&ss->hit, 1u << ss->owner[rb]);
    }
    __syncwarp();
    hit = ss->hit;
    __syncwarp();
  }
  if (pairs) *pairs += (unsigned long long)C;
  return hit;
}

template <int KT, class Store>
__device__ __forceinline__ unsigned long long tx_skip_noops(const GrowArgs& A, Store& st, TxState& t, int64_t seed_i, int lane,
                                                         SkipScratch* ss, unsigned long long budget,
                                                         unsigned long long& iters, bool& halt,
                                                         unsigned long long* dbg = nullptr)
{
  const uint32_t me = (uint32_t)seed_i;
  unsigned long long skipped = 0;
  while (iters + 6 <= budget) {
    const long long c0 = dbg ? clock64() : 0;
    // segment of upcoming entries held by this lane: lane 0 the current frame, lane l the l-th frame below it
    int cur = 0, end = 0;
    if (lane == 0) {
      if (t.have_top) {
        cur = (int)t.top_cur;
        end = (int)t.top_end;
      }
    } else if (t.sp >= lane) {
      const int2 f = st.pop(t.sp - lane);
      cur = f.x;
      end = f.y;
    }
    const uint8_t doomed = ((volatile uint8_t*)A.doom)[seed_i];
    const bool stop = slice_over(A);
    const int len = end - cur;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    const int excl = incl - len;
    const int T = __shfl_sync(FULL_MASK, incl, 31);
    if (T == 0)
      break;
    // upcoming entries `lane` and `lane + 32`: the segment holding each, then the node ids (one round trip)
    int seg0 = 0, seg1 = 0;
#pragma unroll 8
    for (int s = 0; s < 32; ++s) {
      const int v = __shfl_sync(FULL_MASK, incl, s);
      seg0 += (v <= lane) ? 1 : 0;
      seg1 += (v <= lane + 32) ? 1 : 0;
    }
    const bool act0 = lane < T, act1 = lane + 32 < T;
    const int s0 = act0 ? seg0 : 0, s1 = act1 ? seg1 : 0;
    const int e0 = __shfl_sync(FULL_MASK, cur, s0) + (lane - __shfl_sync(FULL_MASK, excl, s0));
    const int e1 = __shfl_sync(FULL_MASK, cur, s1) + (lane + 32 - __shfl_sync(FULL_MASK, excl, s1));
    uint32_t node0 = 0, node1 = 0;
    if (act0) node0 = (uint32_t)st.get(e0);
    if (act1) {  // the second half's rows travel while the first half is evaluated
      node1 = (uint32_t)st.get(e1);
      const int K = KT ? KT : A.K;
      prefetch_l1(A.nbr + (int64_t)node1 * K);
      prefetch_l1(A.nbr + (int64_t)node1 * K + (K - 1));
    }
    const long long c1 = dbg ? clock64() + (long long)((node0 | node1) & 0u) : 0;
    int f;  // entries [0, f) are calls that accept nothing
    uint32_t hit = skip_eval<KT>(A, t.m, me, act0, node0, lane, ss, dbg ? dbg + 4 : nullptr);
    iters += 3;
    if (hit) {
      f = __ffs(hit) - 1;
    } else if (T <= 32) {
      f = T;
    } else {
      hit = skip_eval<KT>(A, t.m, me, act1, node1, lane, ss, dbg ? dbg + 4 : nullptr);
      iters += 2;
      if (dbg && lane == 0) dbg[0] += 1;
      f = 32 + (hit ? __ffs(hit) - 1 : (T < 64 ? T - 32 : 32));
    }
    if (dbg && lane == 0) {
      const long long c3 = clock64() + (long long)(hit & 0u);
      dbg[0] += 1; dbg[1] += (unsigned long long)(c1 - c0); dbg[2] += (unsigned long long)(c3 - c1);
    }
    if (f > 0) {
      skipped += (unsigned long long)f;
      if (f < T) {  // entry f stays: the segment holding it becomes the current frame
        const int sf = __popc(__ballot_sync(FULL_MASK, incl <= f));
        t.top_cur = (int64_t)(__shfl_sync(FULL_MASK, cur, sf) + (f - __shfl_sync(FULL_MASK, excl, sf)));
        t.top_end = (int64_t)__shfl_sync(FULL_MASK, end, sf);
        t.sp -= sf;
      } else {      // every loaded segment is used up
        t.sp -= t.sp < 31 ? t.sp : 31;
        t.top_cur = t.top_end;
      }
      t.have_top = 1;
    }
    if (doomed || stop) {
      halt = true;
      break;
    }
    if (hit || f == 0)
      break;
  }
  return skipped;
}

template <int KT, class Store>
__device__ TxOutcome tx_run_pair(const GrowArgs& A, Store& st, TxState& t, int64_t seed_i, unsigned long long budget,
                                 int lane, unsigned long long& steps_out, unsigned long long* iters_out = nullptr,
                                 SkipScratch* ss = nullptr, unsigned long long* dbg = nullptr)
{
  const int K = KT ? KT : A.K;  // <= 16: one half-warp holds a row
  const uint32_t me = (uint32_t)seed_i;
  const uint32_t fr = A.frontier < 0 ? (uint32_t)__ldcg(A.ctl + CTL_FRONTIER) : (uint32_t)A.frontier;
  const bool fastdiv = (A.flags & GF_FASTDIV) != 0, row_l1 = (A.flags & GF_ROW_L1) != 0;
  const bool skip = ss != nullptr && (A.flags & GF_NOSKIP) == 0;
  const bool approx = (A.flags & GF_EXACT_MODEL) == 0;
  const bool bg = (A.flags & GF_BG) != 0;
  bool halt = false;
  long long klast = dbg ? clock64() : 0;
  const int half = lane >> 4, sl = lane & 15;
  unsigned long long steps = 0, iters = 0;
  TxOutcome out = TX_RUNNING;

  // the node of this lane's half and neighbour column `sl` of it
  bool hasB = !t.depth0 && t.have_top && t.top_cur < t.top_end;
  uint32_t nodeB = hasB ? (uint32_t)st.get(t.top_cur) : 0u;
  int32_t id = -1, stt = 0;
  uint32_t rs = RES_FREE;
  bool mine = false, has_dup = false;
  int4 p = make_int4(0, 0, 0, 0);
  double n0 = 0, n1 = 0, n2 = 0;
  {
    const bool act = half == 0 || hasB;
    const uint32_t node = half ? nodeB : t.node;
    if (act) has_dup = __ldg(A.rowdup + node) != 0;
    if (act && sl >= 1 && sl < K) id = __ldg(A.nbr + (int64_t)node * K + sl);
    if (id >= 0) {
      const int2 sr = ld_state_res(A, id);
      stt = sr.x;
      rs = (uint32_t)sr.y;
      p = __ldg(A.pts + id);
      const double* nr = A.nrm + 3 * (int64_t)id;
      n0 = __ldg(nr); n1 = __ldg(nr + 1); n2 = __ldg(nr + 2);
    }
  }
  while (iters < budget) {  // budget in warp iterations (a skip batch weighs 3): a slice is a TIME slice
    if (!st.reserve((t.len > t.n_at ? t.len : t.n_at) + 3 * K, lane)) {  // before anything is marked
      out = TX_OVERFLOW;
      break;
    }
    ++iters;
    if (dbg && lane == 0) dbg[3] += 1;
    const long long k0 = dbg ? clock64() : 0;
    const bool open = id >= 0 && stt == -1 && rs != me && !mine;
    bool near = false;
    bool want = geo_test_margin(t.m, p, n0, n1, n2, A.th_thick, A.th_dot, near) && open;
    if (!t.model_exact && __any_sync(FULL_MASK, open && near)) {  // too close to call: the reference's own model
      if (fastdiv) model_update_fast(t.m, t.len);
      else model_update(t.m, t.len);
      t.model_exact = 1;
      want = open && geo_test(t.m, p, n0, n1, n2, A.th_thick, A.th_dot);
    }
    const long long k1 = dbg ? clock64() + (long long)(__ballot_sync(FULL_MASK, want) & 0u) : 0;
    if (__any_sync(FULL_MASK, has_dup && want)) {  // a row that names a point twice: first occurrence only
      const unsigned long long key = want ? (((unsigned long long)half << 32) | (uint32_t)id)
                                          : ((1ull << 40) | (unsigned long long)lane);
      const uint32_t same = __match_any_sync(FULL_MASK, key);
      want = want && ((same & lanemask_lt()) == 0);
    }
    if (bg) {
      // Beside the sweeper nothing may be TAKEN from anybody: a slot that is finished and verified is decided by flags, not
      // by a second look at its reservations, and a plane the sweep has accepted is recognised by the reservations it
      // still holds.  A wanted point that a HIGHER in-flight transaction holds (it would be stolen) ends this slot's
      // background slice BEFORE the call is made; the call is made in the next ordinary slice.  A stale reservation
      // (below the frontier of the sweep's start: a transaction decided in an earlier sweep, never one this sweep can
      // accept) is replaced as usual.  What remains is the race with a slot that reserves the same free point at the same
      // moment -- a running slot, which the sweeper cannot decide in this sweep.
      const bool contested = want && rs != RES_FREE && rs > me;
      if (__any_sync(FULL_MASK, contested)) {
        --iters;
        break;
      }
    }
    // ---- node A ----
    bool ok = false, relied = false, fired = false;
    uint32_t old = RES_FREE;
    if (half == 0 && want) {
      ok = true;
      tx_reserve_lane(A, id, rs, me, fr, p.w, ok, relied, fired, old);
    }
    const uint32_t accA = __ballot_sync(FULL_MASK, ok);
    const int cntA = __popc(accA);
    if (t.depth0 && cntA < K - 1) {
      if (ok) st.put(t.len + __popc(accA & lanemask_lt()), id);
      __syncwarp();
      t.len += cntA;
      ++steps;
      out = TX_FAILED;  // :238-239 (under assumptions: the sweeper decides)
      break;
    }
    // ---- node B counts only when A accepted nothing ----
    const bool useB = hasB && cntA == 0;
    if (useB && half == 1 && want) {
      ok = true;
      tx_reserve_lane(A, id, rs, me, fr, p.w, ok, relied, fired, old);
    }
    const uint32_t acc = useB ? __ballot_sync(FULL_MASK, ok) : accA;  // accepted lanes of the LAST call made
    const int cnt = __popc(acc);
    steps += useB ? 2 : 1;
    {  // assumed-taken records of the calls that were made
      const bool rec = relied && (half == 0 || useB);
      const uint32_t rel = __ballot_sync(FULL_MASK, rec);
      if (rel) {
        if (rec) {
          st.put_at(t.n_at + __popc(rel & lanemask_lt()), id);
          atomicMin(A.atby + id, me);
        }
        t.n_at += __popc(rel);
      }
    }
    if (ok) st.put(t.len + __popc(acc & lanemask_lt()), id);
    __syncwarp();
    t.depth0 = 0;
    const long long k2 = dbg ? clock64() + (long long)(acc & 0u) : 0;
    // ---- DFS bookkeeping (:252-255) ----
    if (useB) ++t.top_cur;  // B was the next entry of the frame: consumed
    const int64_t s0 = t.len;
    t.len += cnt;
    uint32_t next = 0, nextB = 0;
    bool have_next = false, have_nextB = false;
    if (cnt > 0) {
      if (t.have_top && t.top_cur < t.top_end) {
        if (lane == 0) st.push(t.sp, make_int2((int)t.top_cur, (int)t.top_end));
        ++t.sp;
      }
      t.top_cur = s0 + 1;  // the first accepted point is visited right away
      t.top_end = t.len;
      t.have_top = 1;
      const int b0 = __ffs(acc) - 1;
      next = (uint32_t)__shfl_sync(FULL_MASK, id, b0);
      have_next = true;
      if (cnt >= 2) {  // the second accepted point follows if the first accepts nothing
        const int b1 = __ffs(acc & (acc - 1)) - 1;
        nextB = (uint32_t)__shfl_sync(FULL_MASK, id, b1);
        have_nextB = true;
      }
    } else {
      // two calls in a row accepted nothing: look at the next 32 nodes at once (tx_skip_noops)
      if (skip && (useB || (A.flags & GF_SKIP_EAGER)) && iters + 6 <= budget) steps += tx_skip_noops<KT>(A, st, t, seed_i, lane, ss, budget, iters, halt, dbg);
      while (t.have_top && t.top_cur == t.top_end) {
        if (t.sp > 0) {
          --t.sp;
          __syncwarp();
          const int2 f = st.pop(t.sp);
          t.top_cur = f.x;
          t.top_end = f.y;
        } else {
          t.have_top = 0;
        }
      }
      if (t.have_top) {
        next = (uint32_t)st.get(t.top_cur);
        ++t.top_cur;
        have_next = true;
        if (t.top_cur < t.top_end) {
          nextB = (uint32_t)st.get(t.top_cur);
          have_nextB = true;
        }
      }
    }
    const long long k3 = dbg ? clock64() + (long long)(next & 0u) : 0;
    // ---- next gathers first, then the model ----
    int32_t nid = -1;
    bool ndup = false;
    {
      const bool act = half == 0 ? have_next : have_nextB;
      const uint32_t node = half ? nextB : next;
      if (act) ndup = __ldg(A.rowdup + node) != 0;
      if (act && sl >= 1 && sl < K) nid = __ldg(A.nbr + (int64_t)node * K + sl);
    }
    const long long k4 = dbg ? clock64() + (long long)(nid & 0) : 0;
    if (cnt > 0) model_accumulate(t.m, acc, p, n0, n1, n2);
    const long long k5 = dbg ? clock64() + (long long)(bseg_d2u(t.m.sn0) & 0ull) : 0;
    int32_t nstt = 0;
    uint32_t nrs = RES_FREE;
    bool nmine = false;
    int4 np = make_int4(0, 0, 0, 0);
    double m0 = 0, m1 = 0, m2 = 0;
    if (nid >= 0) {
      const int2 sr = ld_state_res(A, nid);
      nstt = sr.x;
      nrs = (uint32_t)sr.y;
      np = __ldg(A.pts + nid);
      const double* nr = A.nrm + 3 * (int64_t)nid;
      m0 = __ldg(nr); m1 = __ldg(nr + 1); m2 = __ldg(nr + 2);
      if (row_l1) {  // the row of a neighbour that gets accepted is needed one step later
        prefetch_l1(A.nbr + (int64_t)nid * K);
        prefetch_l1(A.nbr + (int64_t)nid * K + (K - 1));
        prefetch_l1(A.rowdup + nid);
      }
    }
    {  // points accepted a moment ago are ours whatever the gathered reservation says
      uint32_t a = acc;
      while (a) {
        const int b = __ffs(a) - 1;
        a &= a - 1;
        nmine |= nid == __shfl_sync(FULL_MASK, id, b);
      }
    }
    if (cnt > 0) {  // a call that accepts nothing leaves sums and length, hence the model, as they are
      t.model_exact = 0;
      if (!approx || !model_update_approx(t.m, t.len)) {
        if (fastdiv) model_update_fast(t.m, t.len);
        else model_update(t.m, t.len);
        t.model_exact = 1;
      }
    }
    const long long k6 = dbg ? clock64() + (long long)(bseg_d2u(t.m.mn0) & 0ull) : 0;
    {
      bool race = false;
      if (fired) {
        if (old < me && old >= fr) race = true;                  // a lower transaction reserved it in between
        else if (old != RES_FREE && old >= fr) A.doom[old] = 1;  // stolen from a higher transaction: it is void
        if (p.w > (int32_t)me) A.doom[p.w] = 1;                  // the transaction seeded at this point lost its seed
      }
      if (__any_sync(FULL_MASK, race)) {
        out = TX_DOOMED;
        break;
      }
    }
    if (dbg && lane == 0) {
      const long long k7 = clock64();
      dbg[8] += (unsigned long long)(k1 - k0); dbg[9] += (unsigned long long)(k2 - k1); dbg[10] += (unsigned long long)(k3 - k2);
      dbg[11] += (unsigned long long)(k4 - k3); dbg[12] += (unsigned long long)(k5 - k4); dbg[13] += (unsigned long long)(k6 - k5);
      dbg[14] += (unsigned long long)(k7 - k6); dbg[15] += (unsigned long long)(k0 - klast); dbg[16] += cnt > 0 ? 1 : 0;
      klast = k7;
    }
    if (!have_next) {
      out = TX_FINISHED;
      break;
    }
    t.node = next;
    hasB = have_nextB;
    id = nid; stt = nstt; rs = nrs; mine = nmine; p = np; n0 = m0; n1 = m1; n2 = m2;
    has_dup = ndup;
    if ((iters & 7) == 0 || halt) {
      halt = false;
      if (((volatile uint8_t*)A.doom)[seed_i]) {
        out = TX_DOOMED;
        break;
      }
      if (slice_over(A)) break;  // the head finished: let the sweeper commit
    }
  }
  if (out == TX_FINISHED && !t.model_exact) {  // what the plane reports is the reference's model
    if (fastdiv) model_update_fast(t.m, t.len);
    else model_update(t.m, t.len);
    t.model_exact = 1;
  }
  steps_out += steps;
  if (iters_out) *iters_out += iters;
  return out;
}
