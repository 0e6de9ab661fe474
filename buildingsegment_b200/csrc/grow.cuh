// grow.cuh -- types and device helpers shared by the sequential (grow.cu) and speculative
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
  uint8_t* doom;     // 1 = the transaction seeded here overlapped a lower one, or lost its seed
  uint8_t* hasslot;  // 1 = a grower slot is attached to this seed
};

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

// my_function.cpp:241-250 after the accepted points were added to the running sums
__device__ __forceinline__ void model_update(Model& m, int64_t len)
{
  double nn = bseg_sqrt((m.sn0 * m.sn0) + (m.sn1 * m.sn1) + (m.sn2 * m.sn2));
  m.mn0 = m.sn0 / nn; m.mn1 = m.sn1 / nn; m.mn2 = m.sn2 / nn;
  uint64_t dv = (uint64_t)len;
  m.mc0 = (int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)m.sc0) / dv);
  m.mc1 = (int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)m.sc1) / dv);
  m.mc2 = (int32_t)(uint32_t)(((uint64_t)(int64_t)(int32_t)m.sc2) / dv);
}

// add the accepted lanes' normals / positions to the running sums in neighbour order
__device__ __forceinline__ void model_accumulate(Model& m, uint32_t acc, const int4& p, double n0, double n1, double n2)
{
  while (acc) {
    int b = __ffs(acc) - 1;
    acc &= acc - 1;
    m.sn0 += __shfl_sync(FULL_MASK, n0, b);
    m.sn1 += __shfl_sync(FULL_MASK, n1, b);
    m.sn2 += __shfl_sync(FULL_MASK, n2, b);
    m.sc0 += (uint32_t)__shfl_sync(FULL_MASK, p.x, b);
    m.sc1 += (uint32_t)__shfl_sync(FULL_MASK, p.y, b);
    m.sc2 += (uint32_t)__shfl_sync(FULL_MASK, p.z, b);
  }
}

// keep only the lowest lane of every group of accepted lanes that name the same point
// (a row with repeated ids: the second occurrence sees planeIdx already set, :226)
__device__ __forceinline__ bool dedupe(bool ok, int32_t id)
{
  uint32_t same = __match_any_sync(FULL_MASK, ok ? id : -1 - (int)(threadIdx.x & 31));
  return ok && ((same & lanemask_lt()) == 0);
}

