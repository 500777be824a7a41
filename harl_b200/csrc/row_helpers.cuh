// Warp-per-row helpers shared by the head kernels (rowwise.cu) and the trust-region head kernels (trpo.cu).
#pragma once
#include <math.h>

#include "common.cuh"
#include "kernels.cuh"

namespace hb {

constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;
#define HB_LOG_2PI_F 1.8378770664093453f

int tc_dw_splits();
// grid of a gradient kernel that owns one split-buffer slot per CTA
static inline int slot_grid(int64_t rows) {
  int64_t g = ceil_div64(rows, ROW_WARPS);
  int64_t cap = tc_dw_splits() < 296 ? tc_dw_splits() : 296;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

static inline int row_grid(int64_t rows) {
  int64_t g = ceil_div64(rows, ROW_WARPS);
  int64_t cap = 148 * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ------------------------------------------------------------------ heads
template <int HPL>
__device__ __forceinline__ void load_feat(const float* __restrict__ f, int h, int lane, float (&v)[HPL]) {
#pragma unroll
  for (int q = 0; q < HPL; ++q) { int n = lane + 32 * q; v[q] = n < h ? f[n] : 0.f; }
}

// out[j] for j < nout lands in lane j's return value.  The MAXJ dot products are unrolled so that their warp
// reductions interleave (independent shuffle chains) instead of running back to back.
template <int HPL, int MAXJ>
__device__ __forceinline__ float head_linear(const float (&f)[HPL], const float* __restrict__ shw, int h, int nout,
                                             const float* __restrict__ sb, int lane) {
  float mine = 0.f;
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += MAXJ) {
    if (j0 < nout) {
      float p[MAXJ];
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        p[j] = 0.f;
        if (j0 + j < nout) {
#pragma unroll
          for (int q = 0; q < HPL; ++q) { int n = lane + 32 * q; if (n < h) p[j] = fmaf(f[q], shw[(j0 + j) * h + n], p[j]); }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) if (j0 + j < nout) p[j] += __shfl_xor_sync(0xffffffffu, p[j], o);
      }
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) if (lane == j0 + j && j0 + j < nout) mine = p[j] + sb[j0 + j];
    }
  }
  return mine;
}

__device__ __forceinline__ void block_add_scalars(double a, double b, double c, double d, double* out, double* sred) {
  // every lane 0 holds per-warp sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sred[warp * 4 + 0] = a; sred[warp * 4 + 1] = b; sred[warp * 4 + 2] = c; sred[warp * 4 + 3] = d; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int w = 0; w < ROW_WARPS; ++w) s += sred[w * 4 + threadIdx.x];
    atomicAdd(out + threadIdx.x, s);
  }
}

// d(min(s1,s2))/d(ratio) with torch.min / clamp tie semantics (happo.py:71-75)
__device__ __forceinline__ float dmin_dratio(float ratio, float adv, float clip, int use_clip, float* m_out) {
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


// LayerNorm + activation backward of one row held lane-strided by a warp (fused tail of the head kernels).
// df: d loss / d (LN output) for columns lane + 32 q.  Writes dZ to out_row; accumulates dgamma / dbeta partials.
template <int HPL>
__device__ __forceinline__ void ln_act_bwd_row(const float (&df)[HPL], const float (&zrow)[HPL], float mu, float rstd,
                                               const float* __restrict__ lnw, int h, int act, int lane,
                                               float* __restrict__ out_row, float (&cg)[HPL], float (&cb)[HPL]) {
  float g[HPL], xh[HPL], da[HPL];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int q = 0; q < HPL; ++q) {
    const int n = lane + 32 * q;
    g[q] = xh[q] = da[q] = 0.f;
    if (n < h) {
      const float z = zrow[q];
      const float x = (act_fwd_rt(act, z) - mu) * rstd;
      cg[q] = fmaf(df[q], x, cg[q]);
      cb[q] += df[q];
      g[q] = df[q] * lnw[n];
      xh[q] = x;
      da[q] = act_bwd_rt(act, z);
      s1 += g[q];
      s2 = fmaf(g[q], x, s2);
    }
  }
  const float inv_n = 1.f / (float)h;
  const float m1 = warp_sum(s1) * inv_n, m2 = warp_sum(s2) * inv_n;
#pragma unroll
  for (int q = 0; q < HPL; ++q) {
    const int n = lane + 32 * q;
    if (n < h) out_row[n] = rstd * (g[q] - m1 - xh[q] * m2) * da[q];
  }
}

// block-level reduction of the per-lane LN-affine partial sums (all warps) into global memory
template <int HPL>
__device__ __forceinline__ void ln_affine_flush(const float (&cg)[HPL], const float (&cb)[HPL], int h, int lane,
                                                float* sacc /* [2][256] shared, zeroed */, float* g_ln_w, float* g_ln_b,
                                                int64_t slot) {
#pragma unroll
  for (int q = 0; q < HPL; ++q) {
    const int n = lane + 32 * q;
    if (n < h) { atomicAdd(&sacc[n], cg[q]); atomicAdd(&sacc[256 + n], cb[q]); }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < h; n += ROW_THREADS) { acc_out(g_ln_w + n, sacc[n], slot); acc_out(g_ln_b + n, sacc[256 + n], slot); }
}

}  // namespace hb
