// Row-group head kernels for the update phase (evaluate / gradient modes of the Categorical head and the value
// head) at the common trunk widths h in {64, 128}.
//
// The generic kernels in rowwise.cu give one warp to a row; profiling (profiles/ncu_rows_r01_summary.txt) showed
// them instruction-issue bound at ~1000 warp-instructions per row: every dot product of the tiny head pays a
// 5-step warp reduction and all the per-row scalar work (softmax, ratio, clip, LayerNorm statistics) runs once
// per warp-instruction for a single row.  Here LPR = 16 lanes share a row (2 rows per warp), a lane owns CPL = h/16
// columns as float4 chunks, the reductions are 4-step butterflies inside the 16-lane group after which EVERY lane
// holds all logits -- so the softmax / PPO-clip algebra needs no further shuffles -- and each instruction serves
// two rows.  Same maths as rowwise.cu (reference: distributions.py:7-21,37-55; happo.py:66-91; v_critic.py:75-114),
// different summation order inside a row (tolerances in tests/test_gpu_kernels.py).
#include <math.h>

#include "common.cuh"
#include "head_rows.cuh"
#include "kernels.cuh"

namespace hb {

int tc_dw_splits();

namespace {

// threads per CTA: the gradient kernels hold ~240 registers per thread (per-lane head-weight gradient
// accumulators), so a 128-thread CTA lets two of them share an SM and the 296 split-buffer slots stay one wave
constexpr int RT_EVAL = 256, RT_GRAD = 128;
using namespace rows;

template <int ACT>
__device__ __forceinline__ float actf(int rt, float z) {
  if (ACT >= 0) return act_fwd<(ACT >= 0 ? ACT : 0)>(z);
  return act_fwd_rt(rt, z);
}
template <int ACT>
__device__ __forceinline__ float actb(int rt, float z) {
  if (ACT >= 0) return act_bwd<(ACT >= 0 ? ACT : 0)>(z);
  return act_bwd_rt(rt, z);
}


// d(min(s1,s2))/d(ratio) with torch.min / clamp tie semantics (happo.py:71-75)
__device__ __forceinline__ float dmin_dr(float ratio, float adv, float clip, int use_clip, float* m_out) {
  float s1 = ratio * adv;
  if (!use_clip) { *m_out = s1; return adv; }
  float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
  float s2 = rc * adv;
  bool in_range = ratio >= 1.f - clip && ratio <= 1.f + clip;
  *m_out = fminf(s1, s2);
  if (s1 < s2) return adv;
  if (s1 == s2) return in_range ? adv : 0.5f * adv;
  return 0.f;
}
__device__ __forceinline__ float huber_v(float e, float d, int use_huber, float* de) {
  if (!use_huber) { *de = e; return e * e / 2.f; }
  float ae = fabsf(e);
  if (ae <= d) { *de = e; return e * e / 2.f; }
  *de = e > 0.f ? d : -d;
  return d * (ae - d / 2.f);
}

// sum over the LPR lanes of a row group (all lanes of the group end up with the sum)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// sum over the row groups of a warp (lanes with equal lane % LPR)
template <int LPR>
__device__ __forceinline__ float cross_group_sum(float v) {
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

template <int RWARPS>
__device__ __forceinline__ void block_scalars(double a, double b, double c, double d, double* out, double* sred) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum_d(a); b = warp_sum_d(b); c = warp_sum_d(c); d = warp_sum_d(d);
  if (lane == 0) { sred[warp * 4 + 0] = a; sred[warp * 4 + 1] = b; sred[warp * 4 + 2] = c; sred[warp * 4 + 3] = d; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < RWARPS; ++w) s += sred[w * 4 + threadIdx.x];
    atomicAdd(out + threadIdx.x, s);
  }
}

// LayerNorm + activation backward of the CPL columns a lane owns.  df: d loss / d (LN output).
template <int CPL, int LPR, int ACT>
__device__ __forceinline__ void ln_bwd_cols(float (&df)[CPL], const float (&z)[CPL], float mu, float rstd,
                                            const float (&lnw)[CPL], int act_rt, float (&cg)[CPL], float (&cb)[CPL]) {
  float g[CPL], xh[CPL];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const float x = (actf<ACT>(act_rt, z[i]) - mu) * rstd;
    cg[i] = fmaf(df[i], x, cg[i]);
    cb[i] += df[i];
    g[i] = df[i] * lnw[i];
    xh[i] = x;
    s1 += g[i];
    s2 = fmaf(g[i], x, s2);
  }
  const float inv_n = 1.f / (float)(CPL * LPR);
  const float m1 = group_sum<LPR>(s1) * inv_n, m2 = group_sum<LPR>(s2) * inv_n;
#pragma unroll
  for (int i = 0; i < CPL; ++i) df[i] = rstd * (g[i] - m1 - xh[i] * m2) * actb<ACT>(act_rt, z[i]);
}

// ------------------------------------------------------------------ per-warp cp.async row ring
// Rows are streamed through a D-stage shared-memory ring per warp: every lane copies the 16-byte column chunks it
// will consume (feat, and the pre-LN z in gradient mode) with cp.async.cg, and lanes 0..15 of a row group copy that
// row's 4-byte scalars (action, masks, advantage, old log-prob, LN statistics, availability).  D-1 stages (~2 KB
// each per warp) are in flight while one is consumed -- the memory-level parallelism a register prefetch cannot
// afford at ~200 live registers per thread.
constexpr int RING_D = 4;
constexpr int SC_FLOATS = 24;   // per-row scalar block: [0..7] scalars, [8..15] availability, [16..17] int64 source row
enum { SC_ACT = 0, SC_W = 1, SC_FAC = 2, SC_ADV = 3, SC_OLD = 4, SC_REF = 5, SC_MU = 6, SC_RS = 7, SC_AVAIL = 8, SC_SRC = 16 };

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------ Categorical head: evaluate / gradient
template <int CPL, int LPR, int MAXJ, int MODE, int ACT>
__global__ void __launch_bounds__(MODE == MODE_GRAD ? RT_GRAD : RT_EVAL, MODE == MODE_GRAD ? (MAXJ <= 6 ? 3 : 2) : 1) discrete_rows_kernel(HeadArgs a) {
  constexpr int RT = MODE == MODE_GRAD ? RT_GRAD : RT_EVAL, RWARPS = RT / 32;
  constexpr int RPW = 32 / LPR, NC = CPL / 4, H = CPL * LPR;
  constexpr bool GRAD = MODE == MODE_GRAD;
  constexpr int ROW_FLOATS = H * (GRAD ? 2 : 1) + SC_FLOATS;      // feat [+ z] + scalars of one row
  constexpr int STAGE_FLOATS = RPW * ROW_FLOATS;
  extern __shared__ __align__(16) float sm[];
  float* shw = sm;                    // [MAXJ][H]  rows >= na zero
  float* sb = shw + MAXJ * H;         // [8]
  float* sg = sb + 8;                 // [MAXJ][H]  head-weight gradient sums of this CTA
  float* sgb = sg + MAXJ * H;         // [8]
  float* sln = sgb + 8;               // [2][H]     LN affine gradient sums
  double* sred = reinterpret_cast<double*>(sln + 2 * H);          // [RWARPS * 4]
  float* ring = reinterpret_cast<float*>(sred + RWARPS * 4);      // [RWARPS][RING_D][STAGE_FLOATS]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = lane % LPR, rw = lane / LPR;
  const int na = a.out;
  for (int i = threadIdx.x; i < MAXJ * H; i += RT) { shw[i] = i < na * H ? a.hw[i] : 0.f; sg[i] = 0.f; }
  for (int i = threadIdx.x; i < 2 * H; i += RT) sln[i] = 0.f;
  if (threadIdx.x < 8) { sb[threadIdx.x] = threadIdx.x < na ? a.hbias[threadIdx.x] : 0.f; sgb[threadIdx.x] = 0.f; }
  __syncthreads();

  float gacc[GRAD ? MAXJ : 1][CPL];
  float gb[GRAD ? MAXJ : 1];
  float lcg[CPL], lcb[CPL], lnw[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) { lcg[i] = lcb[i] = 0.f; lnw[i] = 1.f; }
  const bool has_ln = GRAD && a.ln_z != nullptr;
  if (GRAD) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      gb[j] = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) gacc[j][i] = 0.f;
    }
    if (has_ln) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float4 v = ld4(a.ln_w + c * 4 * LPR + 4 * s);
        lnw[c * 4 + 0] = v.x; lnw[c * 4 + 1] = v.y; lnw[c * 4 + 2] = v.z; lnw[c * 4 + 3] = v.w;
      }
    }
  }
  const float inv_norm = GRAD ? (float)(1.0 / a.norm3[2]) : 0.f;
  float s_loss = 0.f, s_ent = 0.f, s_ratio = 0.f, s_rows = 0.f;  // per-lane partials over <= ~16 rows; summed in fp64

  // which scalar this lane fetches for its row group: pointer, index multiplier (0: indexed by the batch row r with
  // stride 2 -- the LN statistics; else by the buffer row src) and the slot it lands in
  const float* sc_ptr = nullptr;
  int sc_mul = 1, sc_off = 0, sc_slot = s;
  bool sc_by_r = false;
  if (s == SC_ACT) { if (MODE != MODE_ACT) sc_ptr = a.actions; }
  else if (s == SC_W) { if (GRAD && a.use_active) sc_ptr = a.active; }
  else if (s == SC_FAC) { if (MODE != MODE_ACT) sc_ptr = GRAD ? a.factor : a.factor_inout; }
  else if (s == SC_ADV) { if (GRAD) sc_ptr = a.adv; }
  else if (s == SC_OLD) { if (GRAD) sc_ptr = a.old_logp; }
  else if (s == SC_REF) { if (MODE == MODE_EVAL && a.factor_inout) sc_ptr = a.logp_ref; }
  else if (s == SC_MU || s == SC_RS) { if (has_ln) { sc_ptr = a.ln_stats + (s - SC_MU); sc_by_r = true; sc_mul = 2; } }
  else if (s >= SC_AVAIL && s - SC_AVAIL < na && a.avail != nullptr) { sc_ptr = a.avail; sc_mul = na; sc_off = s - SC_AVAIL; }

  float* wring = ring + (size_t)warp * RING_D * STAGE_FLOATS;
  const int64_t stride = (int64_t)gridDim.x * RWARPS * RPW;
  const int64_t rbase = ((int64_t)blockIdx.x * RWARPS + warp) * RPW + rw;   // this lane's row in iteration 0
  auto src_of = [&](int64_t r) -> int64_t { return (r < a.rows && a.index) ? (int64_t)a.index[r] : r; };
  auto issue = [&](int it, int64_t src) {
    const int64_t r = rbase + (int64_t)it * stride;
    if (r < a.rows) {
      float* row = wring + (it % RING_D) * STAGE_FLOATS + rw * ROW_FLOATS;
#pragma unroll
      for (int c = 0; c < NC; ++c) cp_async16(row + c * 4 * LPR + 4 * s, a.feat + r * H + c * 4 * LPR + 4 * s);
      if (has_ln) {
#pragma unroll
        for (int c = 0; c < NC; ++c) cp_async16(row + H + c * 4 * LPR + 4 * s, a.ln_z + r * H + c * 4 * LPR + 4 * s);
      }
      float* sc = row + H * (GRAD ? 2 : 1);
      if (sc_ptr != nullptr) cp_async4(sc + sc_slot, sc_ptr + (sc_by_r ? r : src) * sc_mul + sc_off);
      if (s == 0) *reinterpret_cast<long long*>(sc + SC_SRC) = (long long)src;
    }
    cp_async_commit();
  };
  // prologue: stages 0 .. D-2
  int64_t q0 = src_of(rbase);
#pragma unroll
  for (int d = 0; d < RING_D - 1; ++d) {
    const int64_t qn = src_of(rbase + (int64_t)(d + 1) * stride);
    issue(d, q0);
    q0 = qn;
  }
  int64_t q1 = src_of(rbase + (int64_t)RING_D * stride);
  int it = 0;
  for (int64_t r0 = rbase - rw; r0 < a.rows; r0 += stride, ++it) {
    const int64_t r = r0 + rw;
    const bool ok = r < a.rows;
    issue(it + RING_D - 1, q0);
    q0 = q1;
    q1 = src_of(r + (int64_t)(RING_D + 1) * stride);
    cp_async_wait<RING_D - 1>();
    __syncwarp();
    const float* row = wring + (it % RING_D) * STAGE_FLOATS + rw * ROW_FLOATS;
    const float* sc = row + H * (GRAD ? 2 : 1);
    float f[CPL];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float4 v = ok ? ld4(row + c * 4 * LPR + 4 * s) : make_float4(0.f, 0.f, 0.f, 0.f);
      f[c * 4 + 0] = v.x; f[c * 4 + 1] = v.y; f[c * 4 + 2] = v.z; f[c * 4 + 3] = v.w;
    }
    // ---- logits: every lane of the group ends up with all of them (head_rows.cuh)
    float lg[MAXJ];
    group_dots<CPL, LPR, MAXJ>(f, shw, s, lg);
    unsigned avm = 0xffu;
    if (a.avail != nullptr && ok) {
      avm = 0u;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) if (j < na && sc[SC_AVAIL + j] != 0.f) avm |= 1u << j;
    }
    float lp[MAXJ], pj[MAXJ];
    const float ent = categorical<MAXJ>(lg, sb, na, avm, lp, pj);
    if constexpr (MODE == MODE_ACT) {
      const int pick = categorical_pick<MAXJ>(pj, na, a.deterministic != 0, a.deterministic ? 0.f : row_uniform(r, a.seed, a.offset + (a.offset_base ? *a.offset_base : 0ull)));
      const float lpp = select<MAXJ>(lp, pick);
      if (ok && s == 0) { a.actions_out[r] = (float)pick; a.logp_out[r] = lpp; }
      __syncwarp();
      continue;
    }
    const int act = ok ? (int)sc[SC_ACT] : 0;
    const float lpa = select<MAXJ>(lp, act);
    if constexpr (MODE == MODE_EVAL) {
      if (ok && s == 0) {
        if (a.logp_out) a.logp_out[r] = lpa;
        if (a.factor_inout) {
          const long long src = *reinterpret_cast<const long long*>(sc + SC_SRC);
          a.factor_inout[src] = sc[SC_FAC] * expf(lpa - sc[SC_REF]);
        }
      }
    } else if constexpr (GRAD) {
      // ---- happo.py:66-91
      float w = 1.f, fac = 1.f, adv = 0.f, old = 0.f;
      if (ok) {
        if (a.use_active) w = sc[SC_W];
        if (a.factor) fac = sc[SC_FAC];
        adv = sc[SC_ADV];
        old = sc[SC_OLD];
      }
      const float ratio = expf(lpa - old);
      float m;
      const float dm = dmin_dr(ratio, adv, a.clip, a.use_clip, &m);
      const float okf = ok ? 1.f : 0.f;
      const float c_lp = -fac * w * inv_norm * dm * ratio * okf;  // d obj / d logp(action)
      const float c_h = a.entropy_coef * w * inv_norm * okf;       // obj has -entropy_coef * H
      if (ok && s == 0) { s_loss += -fac * m * w; s_ent += ent * w; s_ratio += ratio; s_rows += 1.f; }
      float dl[MAXJ];
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const bool live = j < na && ((avm >> j) & 1u);
        dl[j] = live ? c_lp * ((j == act ? 1.f : 0.f) - pj[j]) + c_h * pj[j] * (lp[j] + ent) : 0.f;
        gb[j] += dl[j];
      }
      float df[CPL];
#pragma unroll
      for (int i = 0; i < CPL; ++i) df[i] = 0.f;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float4 w4 = ld4(shw + j * H + c * 4 * LPR + 4 * s);
          df[c * 4 + 0] = fmaf(dl[j], w4.x, df[c * 4 + 0]); df[c * 4 + 1] = fmaf(dl[j], w4.y, df[c * 4 + 1]);
          df[c * 4 + 2] = fmaf(dl[j], w4.z, df[c * 4 + 2]); df[c * 4 + 3] = fmaf(dl[j], w4.w, df[c * 4 + 3]);
        }
#pragma unroll
        for (int i = 0; i < CPL; ++i) gacc[j][i] = fmaf(dl[j], f[i], gacc[j][i]);
      }
      if (has_ln) {
        float z[CPL];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float4 v = ok ? ld4(row + H + c * 4 * LPR + 4 * s) : make_float4(0.f, 0.f, 0.f, 0.f);
          z[c * 4 + 0] = v.x; z[c * 4 + 1] = v.y; z[c * 4 + 2] = v.z; z[c * 4 + 3] = v.w;
        }
        const float ln_mu = ok ? sc[SC_MU] : 0.f, ln_rs = ok ? sc[SC_RS] : 0.f;
        ln_bwd_cols<CPL, LPR, ACT>(df, z, ln_mu, ln_rs, lnw, a.ln_act, lcg, lcb);
      }
      if (ok) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
          st4(a.dfeat + r * H + c * 4 * LPR + 4 * s, make_float4(df[c * 4 + 0], df[c * 4 + 1], df[c * 4 + 2], df[c * 4 + 3]));
      }
    }
    __syncwarp();   // every lane is done with this stage before a later issue overwrites it
  }
  cp_async_wait<0>();
  if constexpr (GRAD) {
    // fold the row groups of the warp, then the warps of the CTA (shared atomics), then one slot write per CTA
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
#pragma unroll
      for (int i = 0; i < CPL; ++i) gacc[j][i] = cross_group_sum<LPR>(gacc[j][i]);
      gb[j] = cross_group_sum<LPR>(gb[j]);
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) { lcg[i] = cross_group_sum<LPR>(lcg[i]); lcb[i] = cross_group_sum<LPR>(lcb[i]); }
    if (rw == 0) {
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        if (j < na) {
#pragma unroll
          for (int i = 0; i < CPL; ++i) atomicAdd(&sg[j * H + (i / 4) * 4 * LPR + 4 * s + (i % 4)], gacc[j][i]);
          if (s == 0) atomicAdd(&sgb[j], gb[j]);
        }
      }
      if (has_ln) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const int col = (i / 4) * 4 * LPR + 4 * s + (i % 4);
          atomicAdd(&sln[col], lcg[i]);
          atomicAdd(&sln[H + col], lcb[i]);
        }
      }
    }
    __syncthreads();
    const int64_t slot = a.part_stride ? a.part_delta + (int64_t)blockIdx.x * a.part_stride : 0;
    for (int i = threadIdx.x; i < na * H; i += RT) acc_out(a.g_hw + i, sg[i], slot);
    if (threadIdx.x < na) acc_out(a.g_hbias + threadIdx.x, sgb[threadIdx.x], slot);
    if (has_ln)
      for (int n = threadIdx.x; n < H; n += RT) { acc_out(a.g_ln_w + n, sln[n], slot); acc_out(a.g_ln_b + n, sln[H + n], slot); }
    block_scalars<RWARPS>((double)s_loss, (double)s_ent, (double)s_ratio, (double)s_rows, a.scalars, sred);
  }
}

// ------------------------------------------------------------------ value head: gradient (v_net.py:65; v_critic.py:75-114)
template <int CPL, int LPR, int ACT>
__global__ void __launch_bounds__(RT_EVAL) value_rows_grad_kernel(ValueArgs a) {
  constexpr int RT = RT_EVAL, RWARPS = RT / 32;
  constexpr int RPW = 32 / LPR, NC = CPL / 4, H = CPL * LPR;
  __shared__ float sgw[H];
  __shared__ float sln[2 * H];
  __shared__ float sgb;
  __shared__ double sred[RWARPS * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = lane % LPR, rw = lane / LPR;
  for (int i = threadIdx.x; i < H; i += RT) { sgw[i] = 0.f; sln[i] = 0.f; sln[H + i] = 0.f; }
  if (threadIdx.x == 0) sgb = 0.f;
  __syncthreads();
  float wv[CPL], gw[CPL], lcg[CPL], lcb[CPL], lnw[CPL];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float4 v = ld4(a.hw + c * 4 * LPR + 4 * s);
    wv[c * 4 + 0] = v.x; wv[c * 4 + 1] = v.y; wv[c * 4 + 2] = v.z; wv[c * 4 + 3] = v.w;
    float4 l = a.ln_z != nullptr ? ld4(a.ln_w + c * 4 * LPR + 4 * s) : make_float4(1.f, 1.f, 1.f, 1.f);
    lnw[c * 4 + 0] = l.x; lnw[c * 4 + 1] = l.y; lnw[c * 4 + 2] = l.z; lnw[c * 4 + 3] = l.w;
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) gw[i] = lcg[i] = lcb[i] = 0.f;
  const float bias = a.hbias[0];
  float vmean = 0.f, vstd = 1.f;
  if (a.vn_state != nullptr) {  // valuenorm.py:38-45
    float d = fmaxf(a.vn_state[2], 1e-5f);
    float mu = a.vn_state[0] / d, msq = a.vn_state[1] / d;
    vmean = mu;
    vstd = sqrtf(fmaxf(msq - mu * mu, 1e-2f));
  }
  float gbias = 0.f, s_loss = 0.f, s_rows = 0.f;
  struct RowIn { float4 f[NC]; float vp, ret; };
  auto fetch = [&](int64_t r, bool ok, int64_t src, RowIn& d) {
#pragma unroll
    for (int c = 0; c < NC; ++c) d.f[c] = ok ? ld4(a.feat + r * H + c * 4 * LPR + 4 * s) : make_float4(0.f, 0.f, 0.f, 0.f);
    d.vp = d.ret = 0.f;
    if (ok) { d.vp = a.value_preds[src]; d.ret = a.returns[src]; }
  };
  const int64_t stride = (int64_t)gridDim.x * RWARPS * RPW;
  int64_t r0 = ((int64_t)blockIdx.x * RWARPS + warp) * RPW;
  auto src_of = [&](int64_t r) -> int64_t { return (r < a.rows && a.index) ? (int64_t)a.index[r] : r; };
  RowIn cur, nxt;
  int64_t src_cur = src_of(r0 + rw), src_nxt = src_of(r0 + stride + rw);
  fetch(r0 + rw, r0 + rw < a.rows, src_cur, cur);
  for (; r0 < a.rows; r0 += stride) {
    const int64_t r = r0 + rw;
    const bool ok = r < a.rows;
    const int64_t src_nn = src_of(r + 2 * stride);
    fetch(r + stride, r + stride < a.rows, src_nxt, nxt);
    float z[CPL];
    float ln_mu = 0.f, ln_rs = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) z[i] = 0.f;
    if (a.ln_z != nullptr && ok) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float4 v = ld4(a.ln_z + r * H + c * 4 * LPR + 4 * s);
        z[c * 4 + 0] = v.x; z[c * 4 + 1] = v.y; z[c * 4 + 2] = v.z; z[c * 4 + 3] = v.w;
      }
      ln_mu = a.ln_stats[r * 2];
      ln_rs = a.ln_stats[r * 2 + 1];
    }
    float f[CPL];
#pragma unroll
    for (int c = 0; c < NC; ++c) { f[c * 4 + 0] = cur.f[c].x; f[c * 4 + 1] = cur.f[c].y; f[c * 4 + 2] = cur.f[c].z; f[c * 4 + 3] = cur.f[c].w; }
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) p = fmaf(f[i], wv[i], p);
    const float v = group_sum<LPR>(p) + bias;
    const float vp = cur.vp;
    float ret = cur.ret;
    if (a.vn_state != nullptr) ret = (ret - vmean) / vstd;
    const float dv = v - vp;
    const float dc = fminf(fmaxf(dv, -a.clip), a.clip);
    const float vclip = vp + dc;
    const bool pass = dv >= -a.clip && dv <= a.clip;
    float de_c, de_o;
    const float l_c = huber_v(ret - vclip, a.huber_delta, a.use_huber, &de_c);
    const float l_o = huber_v(ret - v, a.huber_delta, a.use_huber, &de_o);
    float g_o = -de_o, g_c = pass ? -de_c : 0.f;  // d l / d v
    float loss = l_o, g = g_o;
    if (a.use_clipped) {
      if (l_c > l_o) { loss = l_c; g = g_c; }
      else if (l_c == l_o) { loss = l_o; g = 0.5f * (g_o + g_c); }
    }
    g = ok ? g * a.coef : 0.f;
    if (ok && s == 0) { s_loss += loss; s_rows += 1.f; }
    gbias += g;
    float df[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) { df[i] = g * wv[i]; gw[i] = fmaf(g, f[i], gw[i]); }
    if (a.ln_z != nullptr) ln_bwd_cols<CPL, LPR, ACT>(df, z, ln_mu, ln_rs, lnw, a.ln_act, lcg, lcb);
    if (ok) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        st4(a.dfeat + r * H + c * 4 * LPR + 4 * s, make_float4(df[c * 4 + 0], df[c * 4 + 1], df[c * 4 + 2], df[c * 4 + 3]));
    }
    cur = nxt; src_cur = src_nxt; src_nxt = src_nn;
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) { gw[i] = cross_group_sum<LPR>(gw[i]); lcg[i] = cross_group_sum<LPR>(lcg[i]); lcb[i] = cross_group_sum<LPR>(lcb[i]); }
  gbias = cross_group_sum<LPR>(gbias);
  if (rw == 0) {
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int col = (i / 4) * 4 * LPR + 4 * s + (i % 4);
      atomicAdd(&sgw[col], gw[i]);
      if (a.ln_z != nullptr) { atomicAdd(&sln[col], lcg[i]); atomicAdd(&sln[H + col], lcb[i]); }
    }
    if (s == 0) atomicAdd(&sgb, gbias);
  }
  __syncthreads();
  const int64_t slot = a.part_stride ? a.part_delta + (int64_t)blockIdx.x * a.part_stride : 0;
  for (int n = threadIdx.x; n < H; n += RT) {
    acc_out(a.g_hw + n, sgw[n], slot);
    if (a.ln_z != nullptr) { acc_out(a.g_ln_w + n, sln[n], slot); acc_out(a.g_ln_b + n, sln[H + n], slot); }
  }
  if (threadIdx.x == 0) acc_out(a.g_hbias, sgb, slot);
  block_scalars<RWARPS>((double)s_loss, (double)s_rows, 0.0, 0.0, a.scalars, sred);
}

int grid_for(int64_t rows, int rows_per_cta, bool slots, int slot_cap = 296) {
  int64_t g = ceil_div64(rows, (int64_t)rows_per_cta);
  int64_t cap = slots ? (tc_dw_splits() < slot_cap ? tc_dw_splits() : slot_cap) : 148 * 8;   // gradient kernels: 2-3 CTAs per SM (registers)
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

template <int CPL, int LPR, int MAXJ, int MODE, int ACT>
int launch_discrete(const HeadArgs& a, cudaStream_t st) {
  constexpr int H = CPL * LPR;
  constexpr int RT = MODE == MODE_GRAD ? RT_GRAD : RT_EVAL;
  constexpr int ROW_FLOATS = H * (MODE == MODE_GRAD ? 2 : 1) + SC_FLOATS;
  const size_t smem = (size_t)(2 * MAXJ * H + 16 + 2 * H) * sizeof(float) + (RT / 32) * 4 * sizeof(double) +
                      (size_t)(RT / 32) * RING_D * (32 / LPR) * ROW_FLOATS * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(discrete_rows_kernel<CPL, LPR, MAXJ, MODE, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  auto kern = discrete_rows_kernel<CPL, LPR, MAXJ, MODE, ACT>;
  const int g = grid_for(a.rows, (RT / 32) * (32 / LPR), MODE == MODE_GRAD && a.part_stride != 0, MAXJ <= 6 ? 444 : 296);
  kern<<<g, RT, smem, st>>>(a);
  HB_LAUNCH_DONE(st, shape_label(MODE == MODE_GRAD ? "policy_head_grad" : MODE == MODE_EVAL ? "policy_head_eval" : "policy_head_act", a.rows, a.out, a.h));
  return HB_OK;
}

template <int CPL, int LPR, int MODE, int ACT>
int launch_discrete_j(const HeadArgs& a, cudaStream_t st) {
  if (a.out <= 4) return launch_discrete<CPL, LPR, 4, MODE, ACT>(a, st);
  if (a.out == 5) return launch_discrete<CPL, LPR, 5, MODE, ACT>(a, st);
  if (a.out <= 6) return launch_discrete<CPL, LPR, 6, MODE, ACT>(a, st);
  return launch_discrete<CPL, LPR, 8, MODE, ACT>(a, st);
}

template <int CPL, int LPR, int MODE>
int launch_discrete_act(const HeadArgs& a, cudaStream_t st) {
  // the activation only matters for the fused LayerNorm backward of the gradient mode
  if (MODE == MODE_GRAD && a.ln_z != nullptr && a.ln_act != HB_ACT_RELU) return launch_discrete_j<CPL, LPR, MODE, -1>(a, st);
  return launch_discrete_j<CPL, LPR, MODE, HB_ACT_RELU>(a, st);
}

}  // namespace

// Returns HB_OK and sets *handled when the shape has a row-group kernel; otherwise leaves *handled = false.
int launch_policy_head_rows(int head, int mode, const HeadArgs& a, cudaStream_t st, bool* handled) {
  *handled = false;
  if (head != HB_HEAD_DISCRETE || a.out > 8 || a.rows <= 0) return HB_OK;
  if (a.h != 64 && a.h != 128) return HB_OK;
  *handled = true;
  if (a.h == 64)
    return mode == MODE_GRAD ? launch_discrete_act<4, 16, MODE_GRAD>(a, st)
           : mode == MODE_EVAL ? launch_discrete_act<4, 16, MODE_EVAL>(a, st) : launch_discrete_act<4, 16, MODE_ACT>(a, st);
  return mode == MODE_GRAD ? launch_discrete_act<8, 16, MODE_GRAD>(a, st)
         : mode == MODE_EVAL ? launch_discrete_act<8, 16, MODE_EVAL>(a, st) : launch_discrete_act<8, 16, MODE_ACT>(a, st);
}

int launch_value_head_rows(int grad, const ValueArgs& a, cudaStream_t st, bool* handled) {
  *handled = false;
  if (!grad || a.rows <= 0 || (a.h != 64 && a.h != 128)) return HB_OK;
  *handled = true;
  constexpr int RT = RT_EVAL;
  const int g = grid_for(a.rows, (RT / 32) * 2, a.part_stride != 0);
  const bool relu = a.ln_z == nullptr || a.ln_act == HB_ACT_RELU;
  if (a.h == 64) {
    if (relu) value_rows_grad_kernel<4, 16, HB_ACT_RELU><<<g, RT, 0, st>>>(a);
    else value_rows_grad_kernel<4, 16, -1><<<g, RT, 0, st>>>(a);
  } else {
    if (relu) value_rows_grad_kernel<8, 16, HB_ACT_RELU><<<g, RT, 0, st>>>(a);
    else value_rows_grad_kernel<8, 16, -1><<<g, RT, 0, st>>>(a);
  }
  HB_LAUNCH_DONE(st, shape_label("value_head_grad", a.rows, 1, a.h));
  return HB_OK;
}

}  // namespace hb
