// Row-group arithmetic of the Categorical head, shared by the update-phase kernels (heads_fast.cu: evaluate / gradient
// / act modes) and the fused rollout inference kernel (fused_infer.cu), so that every path computes bit-identical
// log-probabilities and samples for the same inputs.
//
// A row is owned by LPR lanes; lane s holds the columns  c * 4 * LPR + 4 * s + {0..3}  (c < CPL / 4).  After the
// butterfly every lane of the group holds all logits, so softmax / sampling need no further communication.
// Reference: harl/models/base/distributions.py:7-21,37-55 (FixedCategorical over masked logits), act.py:44-80.
#pragma once
#include "common.cuh"

namespace hb {
namespace rows {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Dot products of this lane's CPL columns with the MAXJ head rows (shared memory [MAXJ][CPL * LPR], rows >= n_out
// zero-filled), summed over the LPR lanes of the row group.
template <int CPL, int LPR, int MAXJ>
__device__ __forceinline__ void group_dots(const float (&f)[CPL], const float* __restrict__ shw, int s, float (&lg)[MAXJ]) {
  constexpr int NC = CPL / 4, H = CPL * LPR;
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    float p = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 w4 = ld4(shw + j * H + c * 4 * LPR + 4 * s);
      p = fmaf(f[c * 4 + 0], w4.x, p); p = fmaf(f[c * 4 + 1], w4.y, p);
      p = fmaf(f[c * 4 + 2], w4.z, p); p = fmaf(f[c * 4 + 3], w4.w, p);
    }
    lg[j] = p;
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) lg[j] += __shfl_xor_sync(FULL, lg[j], o);
  }
}

// lg: in = raw dot products, out = masked logits (+bias; -1e10 where unavailable; -inf for j >= na).
// lp = normalised logits (log-probabilities), pj = probabilities, returns the entropy.
template <int MAXJ>
__device__ __forceinline__ float categorical(float (&lg)[MAXJ], const float* __restrict__ sb, int na, unsigned avm,
                                             float (&lp)[MAXJ], float (&pj)[MAXJ]) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    lg[j] = j < na ? (((avm >> j) & 1u) ? lg[j] + sb[j] : -1e10f) : -INFINITY;
    mx = fmaxf(mx, lg[j]);
  }
  float ex[MAXJ];
  float se = 0.f;
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) { ex[j] = expf(lg[j] - mx); se += ex[j]; }
  const float lse = mx + logf(se);
  const float inv_se = 1.f / se;
  float ent = 0.f;
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    lp[j] = lg[j] - lse;
    pj[j] = ex[j] * inv_se;
    ent = fmaf(-fmaxf(lp[j], -3.4028234663852886e38f), pj[j], ent);
  }
  return ent;
}

// mode (first maximum) or inverse-CDF sample with u in (0, 1]; never returns a zero-probability action
template <int MAXJ>
__device__ __forceinline__ int categorical_pick(const float (&pj)[MAXJ], int na, bool deterministic, float u) {
  int act = 0;
  if (deterministic) {
    float pm = -1.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) if (j < na && pj[j] > pm) { pm = pj[j]; act = j; }
    return act;
  }
  float c = 0.f;
  int below = 0;
  unsigned pos = 0u;
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) {
    if (j < na) {
      c += pj[j];
      below += c < u ? 1 : 0;
      pos |= pj[j] > 0.f ? (1u << j) : 0u;
    }
  }
  const int last = 31 - __clz(pos);
  act = below > last ? last : below;
  while (act < 31 && !((pos >> act) & 1u)) ++act;
  return act;
}

template <int MAXJ>
__device__ __forceinline__ float select(const float (&v)[MAXJ], int idx) {
  float r = 0.f;
#pragma unroll
  for (int j = 0; j < MAXJ; ++j) r = j == idx ? v[j] : r;
  return r;
}

// the one uniform a discrete row draws per rollout step (same counter / key layout as the generic head kernels)
__device__ __forceinline__ float row_uniform(long long row, unsigned long long seed, unsigned long long offset) {
  const uint4 rnd = philox4x32(make_uint4((uint32_t)row, (uint32_t)((unsigned long long)row >> 32), 0u, (uint32_t)offset),
                               make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32)));
  return u01(rnd.x);
}

}  // namespace rows
}  // namespace hb
