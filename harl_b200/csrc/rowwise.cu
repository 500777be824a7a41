// Row-wise kernels (one warp per row): feature LayerNorm, LN/activation backward, and the
// three heads (Categorical, DiagGaussian, value) in act / evaluate / gradient modes.
// These are HBM/L2-bandwidth-bound: coalesced lane-strided row accesses, warp-shuffle
// reductions, register accumulation of per-column gradient sums, one shared-memory and one
// global atomic pass per CTA at the end.
#include <math.h>

#include "common.cuh"
#include "kernels.cuh"
#include "row_helpers.cuh"

namespace hb {

// ------------------------------------------------------------------ feature LayerNorm (mlp.py:57-58,65-66)
// xout[r][0:kpad] = normalised (no affine: folded into layer 0 by hb_net_prepare) or raw copy; zero padded.
__global__ void __launch_bounds__(ROW_THREADS) feat_norm_kernel(const float* __restrict__ obs, int in_dim,
                                                                const int32_t* __restrict__ index, int64_t rows,
                                                                int feature_norm, float* __restrict__ xout, int ldx) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * ROW_WARPS;
  for (int64_t r = w0; r < rows; r += nw) {
    const int64_t src = index ? (int64_t)index[r] : r;
    const float* o = obs + src * in_dim;
    float mean = 0.f, rstd = 1.f;
    if (feature_norm) {
      float s = 0.f;
      for (int k = lane; k < in_dim; k += 32) s += o[k];
      mean = warp_sum(s) / (float)in_dim;
      float q = 0.f;
      for (int k = lane; k < in_dim; k += 32) { float dlt = o[k] - mean; q = fmaf(dlt, dlt, q); }
      rstd = rsqrtf(warp_sum(q) / (float)in_dim + 1e-5f);
    }
    float* x = xout + r * ldx;
    for (int k = lane; k < ldx; k += 32) x[k] = k < in_dim ? (o[k] - mean) * rstd : 0.f;
  }
}

// Narrow rows (in_dim <= 44: MPE / MAMuJoCo observations): one thread per row.  A CTA stages 256 rows through
// shared memory with fully coalesced global loads / stores (row pitch in_dim+1 -> conflict-free per-thread walks).
__global__ void __launch_bounds__(256) feat_norm_narrow_kernel(const float* __restrict__ obs, int in_dim,
                                                               const int32_t* __restrict__ index, int64_t rows,
                                                               int feature_norm, float* __restrict__ xout, int ldx) {
  extern __shared__ float srow[];
  const int pitch = (in_dim > ldx ? in_dim : ldx) + 1;
  const int64_t r0 = (int64_t)blockIdx.x * 256;
  const int nrows = (int)(rows - r0 < 256 ? rows - r0 : 256);
  if (index == nullptr) {
    const float* base = obs + r0 * in_dim;
    for (int f = threadIdx.x; f < nrows * in_dim; f += 256) srow[(f / in_dim) * pitch + f % in_dim] = base[f];
  } else {
    for (int f = threadIdx.x; f < nrows * in_dim; f += 256) {
      int r = f / in_dim, k = f % in_dim;
      srow[r * pitch + k] = obs[(int64_t)index[r0 + r] * in_dim + k];
    }
  }
  __syncthreads();
  if (threadIdx.x < nrows) {
    float* o = srow + threadIdx.x * pitch;
    float mean = 0.f, rstd = 1.f;
    if (feature_norm) {
      float s = 0.f;
      for (int k = 0; k < in_dim; ++k) s += o[k];
      mean = s / (float)in_dim;
      float q = 0.f;
      for (int k = 0; k < in_dim; ++k) { float dlt = o[k] - mean; q = fmaf(dlt, dlt, q); }
      rstd = rsqrtf(q / (float)in_dim + 1e-5f);
    }
    for (int k = 0; k < ldx; ++k) o[k] = k < in_dim ? (o[k] - mean) * rstd : 0.f;
  }
  __syncthreads();
  float* out = xout + r0 * ldx;
  for (int f = threadIdx.x; f < nrows * ldx; f += 256) out[f] = srow[(f / ldx) * pitch + f % ldx];
}

int launch_feat_norm(const float* obs, int in_dim, const int32_t* index, int64_t rows, int feature_norm, float* xout,
                     int ldx, cudaStream_t st) {
  if (rows <= 0) return HB_OK;
  if (in_dim <= 44) {
    size_t smem = (size_t)256 * ((in_dim > ldx ? in_dim : ldx) + 1) * sizeof(float);
    feat_norm_narrow_kernel<<<(unsigned)ceil_div64(rows, 256), 256, smem, st>>>(obs, in_dim, index, rows, feature_norm,
                                                                              xout, ldx);
    HB_LAUNCH_DONE(st, shape_label("feat_norm", rows, ldx, in_dim));
    return HB_OK;
  }
  feat_norm_kernel<<<row_grid(rows), ROW_THREADS, 0, st>>>(obs, in_dim, index, rows, feature_norm, xout, ldx);
  HB_LAUNCH_DONE(st, shape_label("feat_norm", rows, ldx, in_dim));
  return HB_OK;
}

// ------------------------------------------------------------------ LN + activation backward of the last trunk block
// dZ = act'(Z) * LNbwd(dY);  g_lnw += sum_r dY*xhat;  g_lnb += sum_r dY.   (dY and dZ may alias)
__global__ void __launch_bounds__(ROW_THREADS) ln_act_bwd_kernel(const float* dY, const float* __restrict__ Z,
                                                                 const float* __restrict__ stats,
                                                                 const float* __restrict__ lnw, float* dZ,
                                                                 float* __restrict__ g_lnw, float* __restrict__ g_lnb,
                                                                 int64_t rows, int N, int act) {
  __shared__ float red[ROW_WARPS][2][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t w0 = (int64_t)blockIdx.x * ROW_WARPS + warp, nw = (int64_t)gridDim.x * ROW_WARPS;
  float cg[8], cb[8], gam[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { cg[q] = cb[q] = 0.f; int n = lane + 32 * q; gam[q] = n < N ? lnw[n] : 0.f; }
  const float inv_n = 1.f / (float)N;
  for (int64_t r = w0; r < rows; r += nw) {
    const float mu = stats[r * 2], rstd = stats[r * 2 + 1];
    float g[8], xh[8], da[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int n = lane + 32 * q;
      g[q] = xh[q] = da[q] = 0.f;
      if (n < N) {
        float dy = dY[r * N + n], z = Z[r * N + n];
        float x = (act_fwd_rt(act, z) - mu) * rstd;
        cg[q] = fmaf(dy, x, cg[q]);
        cb[q] += dy;
        g[q] = dy * gam[q];
        xh[q] = x;
        da[q] = act_bwd_rt(act, z);
        s1 += g[q];
        s2 = fmaf(g[q], x, s2);
      }
    }
    const float m1 = warp_sum(s1) * inv_n, m2 = warp_sum(s2) * inv_n;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int n = lane + 32 * q;
      if (n < N) dZ[r * N + n] = rstd * (g[q] - m1 - xh[q] * m2) * da[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) { red[warp][0][lane + 32 * q] = cg[q]; red[warp][1][lane + 32 * q] = cb[q]; }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += ROW_THREADS) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < ROW_WARPS; ++w) { a += red[w][0][n]; b += red[w][1][n]; }
    atomicAdd(g_lnw + n, a);
    atomicAdd(g_lnb + n, b);
  }
}

int launch_ln_act_bwd(const float* dY, const float* Z, const float* stats, const float* lnw, float* dZ, float* g_lnw,
                      float* g_lnb, int64_t rows, int N, int act, cudaStream_t st) {
  if (rows <= 0) return HB_OK;
  ln_act_bwd_kernel<<<row_grid(rows), ROW_THREADS, 0, st>>>(dY, Z, stats, lnw, dZ, g_lnw, g_lnb, rows, N, act);
  HB_LAUNCH_DONE(st, shape_label("ln_act_bwd", rows, N, 0));
  return HB_OK;
}

// ---- Categorical (distributions.py:7-21,37-55; act.py:44-80,143-155)
template <int HPL, int MAXJ, int MODE>
__global__ void __launch_bounds__(ROW_THREADS) discrete_head_kernel(HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* shw = sm;                      // [out][h]
  float* sb = shw + a.out * a.h;        // [out]
  float* sg = sb + 32;                  // [out][h] grad accumulation (MODE_GRAD)
  float* sgb = sg + (MODE == MODE_GRAD ? a.out * a.h : 0);  // [32]
  double* sred = reinterpret_cast<double*>(sgb + 32);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = a.h, na = a.out;
  for (int i = threadIdx.x; i < na * h; i += ROW_THREADS) { shw[i] = a.hw[i]; if (MODE == MODE_GRAD) sg[i] = 0.f; }
  if (threadIdx.x < 32) { sb[threadIdx.x] = threadIdx.x < na ? a.hbias[threadIdx.x] : 0.f; sgb[threadIdx.x] = 0.f; }
  __syncthreads();
  const int64_t w0 = (int64_t)blockIdx.x * ROW_WARPS + warp, nw = (int64_t)gridDim.x * ROW_WARPS;
  float gacc[MODE == MODE_GRAD ? MAXJ : 1][HPL];
  float gb = 0.f;
  float lcg[HPL], lcb[HPL];
#pragma unroll
  for (int q = 0; q < HPL; ++q) lcg[q] = lcb[q] = 0.f;
  __shared__ float s_ln[MODE == MODE_GRAD ? 512 : 1];
  if (MODE == MODE_GRAD) { for (int i = threadIdx.x; i < 512; i += ROW_THREADS) s_ln[i] = 0.f; }
  double s_loss = 0.0, s_ent = 0.0, s_ratio = 0.0, s_rows = 0.0;
  if (MODE == MODE_GRAD) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
#pragma unroll
      for (int q = 0; q < HPL; ++q) gacc[j][q] = 0.f;
  }
  const float inv_norm = MODE == MODE_GRAD ? (float)(1.0 / a.norm3[2]) : 0.f;
  // Software pipeline: the loads of row r + nw (and the gather index of row r + 2 nw) are in flight while row r is
  // computed, so a warp pays the two dependent memory round trips (index -> row scalars) once, not once per row.
  struct RowIn { float f[HPL], zr[HPL]; float mu, rs, act, w, fac, adv, old, av, ref; };
  auto fetch = [&](int64_t r, int64_t src, RowIn& d) {
    load_feat<HPL>(a.feat + r * h, h, lane, d.f);
    d.av = 1.f;
    if (lane < na && a.avail != nullptr) d.av = a.avail[src * na + lane];
#pragma unroll
    for (int q = 0; q < HPL; ++q) d.zr[q] = 0.f;
    d.mu = d.rs = d.act = d.adv = d.old = d.ref = 0.f;
    d.w = d.fac = 1.f;
    if constexpr (MODE == MODE_GRAD) {
      if (a.ln_z != nullptr) { load_feat<HPL>(a.ln_z + r * h, h, lane, d.zr); d.mu = a.ln_stats[r * 2]; d.rs = a.ln_stats[r * 2 + 1]; }
      d.act = a.actions[src];
      if (a.use_active) d.w = a.active[src];
      if (a.factor) d.fac = a.factor[src];
      d.adv = a.adv[src];
      d.old = a.old_logp[src];
    }
    if constexpr (MODE == MODE_EVAL) {
      d.act = a.actions[src];
      if (a.factor_inout) { d.ref = a.logp_ref[src]; d.fac = a.factor_inout[src]; }
    }
  };
  RowIn cur, nxt;
  int64_t src_cur = 0, src_nxt = 0;
  if (w0 < a.rows) { src_cur = a.index ? (int64_t)a.index[w0] : w0; fetch(w0, src_cur, cur); }
  if (w0 + nw < a.rows) src_nxt = a.index ? (int64_t)a.index[w0 + nw] : w0 + nw;
  for (int64_t r = w0; r < a.rows; r += nw) {
    const int64_t src = src_cur;
    int64_t src_nn = 0;
    if (r + 2 * nw < a.rows) src_nn = a.index ? (int64_t)a.index[r + 2 * nw] : r + 2 * nw;
    if (r + nw < a.rows) fetch(r + nw, src_nxt, nxt);
    float f[HPL], zr[HPL];
#pragma unroll
    for (int q = 0; q < HPL; ++q) { f[q] = cur.f[q]; zr[q] = cur.zr[q]; }
    const float ln_mu = cur.mu, ln_rs = cur.rs, sc_act = cur.act, sc_w = cur.w, sc_fac = cur.fac, sc_adv = cur.adv, sc_old = cur.old;
    const float sc_ref = cur.ref;
    const bool valid = lane < na;
    const bool masked = valid && cur.av == 0.f;
    (void)zr; (void)ln_mu; (void)ln_rs; (void)sc_w; (void)sc_adv; (void)sc_old; (void)sc_ref; (void)src;
    float logit = head_linear<HPL, (MAXJ < 8 ? MAXJ : 8)>(f, shw, h, na, sb, lane);
    if (masked) logit = -1e10f;
    const float mx = warp_max(valid ? logit : -INFINITY);
    const float ex = valid ? expf(logit - mx) : 0.f;
    const float lse = mx + logf(warp_sum(ex));
    const float lp = valid ? logit - lse : 0.f;  // normalised logit (torch Categorical(logits=))
    const float p = valid ? expf(lp) : 0.f;
    if constexpr (MODE == MODE_ACT) {
      int act;
      if (a.deterministic) {
        float pm = warp_max(p);
        unsigned b = __ballot_sync(0xffffffffu, valid && p == pm);
        act = __ffs(b) - 1;
      } else {
        const uint64_t off = a.offset + (a.offset_base ? *a.offset_base : 0ull);
        uint4 rnd = philox4x32(make_uint4((uint32_t)r, (uint32_t)((uint64_t)r >> 32), 0u, (uint32_t)off),
                               make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ (uint32_t)(off >> 32)));
        float u = u01(rnd.x);
        float c = p;  // inclusive prefix sum
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, c, o); if (lane >= o) c += t; }
        unsigned below = __ballot_sync(0xffffffffu, valid && c < u);
        unsigned pos = __ballot_sync(0xffffffffu, valid && p > 0.f);
        int last = 31 - __clz(pos);
        act = __popc(below);
        if (act > last) act = last;
        // never return a zero-probability action
        while (act < 31 && !((pos >> act) & 1u)) ++act;
      }
      float lpa = __shfl_sync(0xffffffffu, lp, act);
      if (lane == 0) { a.actions_out[r] = (float)act; a.logp_out[r] = lpa; }
    } else if constexpr (MODE == MODE_EVAL) {
      const int act = (int)sc_act;
      const float lpa = __shfl_sync(0xffffffffu, lp, act);
      if (lane == 0) {
        if (a.logp_out) a.logp_out[r] = lpa;
        if (a.factor_inout) a.factor_inout[src] = sc_fac * expf(lpa - sc_ref);
      }
    } else {
    // ---- MODE_GRAD: happo.py:66-91
    const int act = (int)sc_act;
    const float lpa = __shfl_sync(0xffffffffu, lp, act);
    const float ent = -warp_sum(valid ? fmaxf(lp, -3.4028234663852886e38f) * p : 0.f);
    const float w = sc_w, fac = sc_fac, adv = sc_adv;
    const float ratio = expf(lpa - sc_old);
    float m;
    const float dm = dmin_dratio(ratio, adv, a.clip, a.use_clip, &m);
    const float c_lp = -fac * w * inv_norm * dm * ratio;       // d obj / d logp(action)
    const float c_h = a.entropy_coef * w * inv_norm;             // obj has -entropy_coef * H
    float dl = 0.f;
    if (valid && !masked) dl = c_lp * ((lane == act ? 1.f : 0.f) - p) + c_h * p * (lp + ent);
    if (lane == 0) { s_loss += (double)(-fac * m * w); s_ent += (double)(ent * w); s_ratio += (double)ratio; s_rows += 1.0; }
    gb += dl;
    float df[HPL];
#pragma unroll
    for (int q = 0; q < HPL; ++q) df[q] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < na) {
        const float dj = __shfl_sync(0xffffffffu, dl, j);
#pragma unroll
        for (int q = 0; q < HPL; ++q) {
          int n = lane + 32 * q;
          if (n < h) { df[q] = fmaf(dj, shw[j * h + n], df[q]); gacc[j][q] = fmaf(dj, f[q], gacc[j][q]); }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < HPL; ++q) { int n = lane + 32 * q; if (n < h && a.ln_z == nullptr) a.dfeat[r * h + n] = df[q]; }
    if (a.ln_z != nullptr)
      ln_act_bwd_row<HPL>(df, zr, ln_mu, ln_rs, a.ln_w, h, a.ln_act, lane, a.dfeat + r * h, lcg, lcb);
    }
    cur = nxt; src_cur = src_nxt; src_nxt = src_nn;
  }
  if constexpr (MODE == MODE_GRAD) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (j < na) {
#pragma unroll
        for (int q = 0; q < HPL; ++q) { int n = lane + 32 * q; if (n < h) atomicAdd(&sg[j * h + n], gacc[j][q]); }
      }
    if (lane < na) atomicAdd(&sgb[lane], gb);
    __syncthreads();
    const int64_t slot = a.part_stride ? a.part_delta + (int64_t)blockIdx.x * a.part_stride : 0;
    for (int i = threadIdx.x; i < na * h; i += ROW_THREADS) acc_out(a.g_hw + i, sg[i], slot);
    if (threadIdx.x < na) acc_out(a.g_hbias + threadIdx.x, sgb[threadIdx.x], slot);
    block_add_scalars(s_loss, s_ent, s_ratio, s_rows, a.scalars, sred);
    if (a.ln_z != nullptr) ln_affine_flush<HPL>(lcg, lcb, h, lane, s_ln, a.g_ln_w, a.g_ln_b, slot);
  }
}

// ---- DiagGaussian (distributions.py:24-34,58-89)
template <int HPL, int MAXJ, int MODE>
__global__ void __launch_bounds__(ROW_THREADS) box_head_kernel(HeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* shw = sm;
  float* sb = shw + a.out * a.h;
  float* sg = sb + 32;
  float* sgb = sg + (MODE == MODE_GRAD ? a.out * a.h : 0);  // [32] bias grads, then [32] log_std grads
  double* sred = reinterpret_cast<double*>(sgb + 64);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = a.h, ad = a.out;
  for (int i = threadIdx.x; i < ad * h; i += ROW_THREADS) { shw[i] = a.hw[i]; if (MODE == MODE_GRAD) sg[i] = 0.f; }
  if (threadIdx.x < 32) sb[threadIdx.x] = threadIdx.x < ad ? a.hbias[threadIdx.x] : 0.f;
  if (threadIdx.x < 64) sgb[threadIdx.x] = 0.f;
  __syncthreads();
  const bool valid = lane < ad;
  float sig = 0.f, std = 1.f;
  if (valid) { sig = 1.f / (1.f + expf(-a.log_std[lane] / a.std_x)); std = sig * a.std_y; }
  const float log_std_v = logf(std);
  const float dstd_ds = a.std_y * sig * (1.f - sig) / a.std_x;  // d std / d log_std param
  const float ent_row = warp_sum(valid ? 0.5f + 0.5f * HB_LOG_2PI_F + log_std_v : 0.f);
  const int64_t w0 = (int64_t)blockIdx.x * ROW_WARPS + warp, nw = (int64_t)gridDim.x * ROW_WARPS;
  float gacc[MODE == MODE_GRAD ? MAXJ : 1][HPL];
  float gb = 0.f, gs = 0.f;
  float lcg[HPL], lcb[HPL];
#pragma unroll
  for (int q = 0; q < HPL; ++q) lcg[q] = lcb[q] = 0.f;
  __shared__ float s_ln[MODE == MODE_GRAD ? 512 : 1];
  if (MODE == MODE_GRAD) { for (int i = threadIdx.x; i < 512; i += ROW_THREADS) s_ln[i] = 0.f; }
  double s_loss = 0.0, s_ent = 0.0, s_ratio = 0.0, s_rows = 0.0;
  if (MODE == MODE_GRAD) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
#pragma unroll
      for (int q = 0; q < HPL; ++q) gacc[j][q] = 0.f;
  }
  const float inv_norm = MODE == MODE_GRAD ? (float)(1.0 / a.norm3[2]) : 0.f;
  // software pipeline as in discrete_head_kernel
  struct RowIn { float f[HPL], zr[HPL]; float mu, rs, act, old, w, fac, adv, ref; };
  auto fetch = [&](int64_t r, int64_t src, RowIn& d) {
    load_feat<HPL>(a.feat + r * h, h, lane, d.f);
#pragma unroll
    for (int q = 0; q < HPL; ++q) d.zr[q] = 0.f;
    d.mu = d.rs = d.act = d.old = d.adv = d.ref = 0.f;
    d.w = d.fac = 1.f;
    if constexpr (MODE != MODE_ACT) { if (valid) d.act = a.actions[src * ad + lane]; }
    if constexpr (MODE == MODE_EVAL) {
      if (a.factor_inout) { if (valid) d.ref = a.logp_ref[src * ad + lane]; d.fac = a.factor_inout[src]; }
    }
    if constexpr (MODE == MODE_GRAD) {
      if (a.ln_z != nullptr) { load_feat<HPL>(a.ln_z + r * h, h, lane, d.zr); d.mu = a.ln_stats[r * 2]; d.rs = a.ln_stats[r * 2 + 1]; }
      if (valid) d.old = a.old_logp[src * ad + lane];
      if (a.use_active) d.w = a.active[src];
      if (a.factor) d.fac = a.factor[src];
      d.adv = a.adv[src];
    }
  };
  RowIn cur, nxt;
  int64_t src_cur = 0, src_nxt = 0;
  if (w0 < a.rows) { src_cur = a.index ? (int64_t)a.index[w0] : w0; fetch(w0, src_cur, cur); }
  if (w0 + nw < a.rows) src_nxt = a.index ? (int64_t)a.index[w0 + nw] : w0 + nw;
  for (int64_t r = w0; r < a.rows; r += nw) {
    const int64_t src = src_cur;
    int64_t src_nn = 0;
    if (r + 2 * nw < a.rows) src_nn = a.index ? (int64_t)a.index[r + 2 * nw] : r + 2 * nw;
    if (r + nw < a.rows) fetch(r + nw, src_nxt, nxt);
    float f[HPL], zr[HPL];
#pragma unroll
    for (int q = 0; q < HPL; ++q) { f[q] = cur.f[q]; zr[q] = cur.zr[q]; }
    const float ln_mu = cur.mu, ln_rs = cur.rs;
    (void)zr; (void)ln_mu; (void)ln_rs; (void)src;
    const float mean = head_linear<HPL, (MAXJ < 8 ? MAXJ : 8)>(f, shw, h, ad, sb, lane);
    if constexpr (MODE == MODE_ACT) {
      float act = mean;
      if (!a.deterministic) {
        const uint64_t off = a.offset + (a.offset_base ? *a.offset_base : 0ull);
        uint4 rnd = philox4x32(make_uint4((uint32_t)r, (uint32_t)((uint64_t)r >> 32), (uint32_t)lane, (uint32_t)off),
                               make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ (uint32_t)(off >> 32)));
        float u1 = u01(rnd.x), u2 = u01(rnd.y);
        float z = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
        act = mean + std * z;
      }
      if (valid) {
        float dlt = act - mean;
        a.actions_out[r * ad + lane] = act;
        a.logp_out[r * ad + lane] = -(dlt * dlt) / (2.f * std * std) - log_std_v - 0.5f * HB_LOG_2PI_F;
      }
    } else {
    const float act = cur.act;
    const float dlt = act - mean;
    const float var = std * std;
    const float lp = -(dlt * dlt) / (2.f * var) - log_std_v - 0.5f * HB_LOG_2PI_F;
    if constexpr (MODE == MODE_EVAL) {
      if (valid && a.logp_out) a.logp_out[r * ad + lane] = lp;
      if (a.factor_inout) {
        float e = valid ? expf(lp - cur.ref) : (a.agg_prod ? 1.f : 0.f);
        float agg;
        if (a.agg_prod) {
          agg = e;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) agg *= __shfl_xor_sync(0xffffffffu, agg, o);
        } else {
          agg = warp_sum(e) / (float)ad;
        }
        if (lane == 0) a.factor_inout[src] = cur.fac * agg;
      }
    } else {
    // ---- MODE_GRAD
    const float e = valid ? expf(lp - cur.old) : (a.agg_prod ? 1.f : 0.f);
    float ratio;
    if (a.agg_prod) {
      ratio = e;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ratio *= __shfl_xor_sync(0xffffffffu, ratio, o);
    } else {
      ratio = warp_sum(e) / (float)ad;
    }
    const float w = cur.w, fac = cur.fac, adv = cur.adv;
    float m;
    const float dm = dmin_dratio(ratio, adv, a.clip, a.use_clip, &m);
    const float c_r = -fac * w * inv_norm * dm;   // d obj / d ratio
    // d ratio / d lp_d: prod -> ratio (= prod_k e_k; d/de_d * e_d), mean -> e_d / ad
    const float dr_dlp = a.agg_prod ? ratio : e / (float)ad;
    const float c_lp = valid ? c_r * dr_dlp : 0.f;
    const float dmean = c_lp * dlt / var;
    const float dstd = c_lp * (dlt * dlt / (var * std) - 1.f / std) - a.entropy_coef * w * inv_norm / std;
    if (valid) gs += dstd * dstd_ds;
    if (lane == 0) { s_loss += (double)(-fac * m * w); s_ent += (double)(ent_row * w); s_ratio += (double)ratio; s_rows += 1.0; }
    gb += dmean;
    float df[HPL];
#pragma unroll
    for (int q = 0; q < HPL; ++q) df[q] = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      if (j < ad) {
        const float dj = __shfl_sync(0xffffffffu, dmean, j);
#pragma unroll
        for (int q = 0; q < HPL; ++q) {
          int n = lane + 32 * q;
          if (n < h) { df[q] = fmaf(dj, shw[j * h + n], df[q]); gacc[j][q] = fmaf(dj, f[q], gacc[j][q]); }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < HPL; ++q) { int n = lane + 32 * q; if (n < h && a.ln_z == nullptr) a.dfeat[r * h + n] = df[q]; }
    if (a.ln_z != nullptr)
      ln_act_bwd_row<HPL>(df, zr, ln_mu, ln_rs, a.ln_w, h, a.ln_act, lane, a.dfeat + r * h, lcg, lcb);
    }
    }
    cur = nxt; src_cur = src_nxt; src_nxt = src_nn;
  }
  if constexpr (MODE == MODE_GRAD) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
      if (j < ad) {
#pragma unroll
        for (int q = 0; q < HPL; ++q) { int n = lane + 32 * q; if (n < h) atomicAdd(&sg[j * h + n], gacc[j][q]); }
      }
    if (valid) { atomicAdd(&sgb[lane], gb); atomicAdd(&sgb[32 + lane], gs); }
    __syncthreads();
    const int64_t slot = a.part_stride ? a.part_delta + (int64_t)blockIdx.x * a.part_stride : 0;
    for (int i = threadIdx.x; i < ad * h; i += ROW_THREADS) acc_out(a.g_hw + i, sg[i], slot);
    if (threadIdx.x < ad) { acc_out(a.g_hbias + threadIdx.x, sgb[threadIdx.x], slot); acc_out(a.g_log_std + threadIdx.x, sgb[32 + threadIdx.x], slot); }
    block_add_scalars(s_loss, s_ent, s_ratio, s_rows, a.scalars, sred);
    if (a.ln_z != nullptr) ln_affine_flush<HPL>(lcg, lcb, h, lane, s_ln, a.g_ln_w, a.g_ln_b, slot);
  }
}

#define HB_HEAD_LAUNCH(KERN, HPLV, MAXJV, MODEV)                                                              \
  do {                                                                                                        \
    auto kern = KERN<HPLV, MAXJV, MODEV>;                                                                     \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    kern<<<(MODEV == MODE_GRAD && a.part_stride) ? slot_grid(a.rows) : row_grid(a.rows), ROW_THREADS, smem, st>>>(a);                                                     \
  } while (0)

template <int MODE>
static int launch_head_mode(int head, const HeadArgs& a, cudaStream_t st) {
  if (a.rows <= 0) return HB_OK;
  const int hpl = a.h <= 32 ? 1 : a.h <= 64 ? 2 : a.h <= 128 ? 4 : 8;
  int maxj = 8;
  if (MODE == MODE_GRAD) {
    maxj = a.out <= 8 ? 8 : a.out <= 16 ? 16 : 32;
    if (maxj * hpl > 64) {
      set_error("head gradient: out_dim %d with hidden %d exceeds the register-accumulator budget", a.out, a.h);
      return HB_ERR_UNSUPPORTED;
    }
  }
  size_t smem = (size_t)(a.out * a.h + 32) * 4 + (MODE == MODE_GRAD ? (size_t)a.out * a.h * 4 : 0) + 64 * 4 +
                ROW_WARPS * 4 * sizeof(double) + 16;
#define HB_HEAD_DISPATCH(KERN)                                                                      \
  switch (hpl * 100 + maxj) {                                                                       \
    case 108: HB_HEAD_LAUNCH(KERN, 1, 8, MODE); break;                                              \
    case 116: HB_HEAD_LAUNCH(KERN, 1, 16, MODE); break;                                             \
    case 132: HB_HEAD_LAUNCH(KERN, 1, 32, MODE); break;                                             \
    case 208: HB_HEAD_LAUNCH(KERN, 2, 8, MODE); break;                                              \
    case 216: HB_HEAD_LAUNCH(KERN, 2, 16, MODE); break;                                             \
    case 232: HB_HEAD_LAUNCH(KERN, 2, 32, MODE); break;                                             \
    case 408: HB_HEAD_LAUNCH(KERN, 4, 8, MODE); break;                                              \
    case 416: HB_HEAD_LAUNCH(KERN, 4, 16, MODE); break;                                             \
    case 808: HB_HEAD_LAUNCH(KERN, 8, 8, MODE); break;                                              \
    default: set_error("head dispatch %d/%d", hpl, maxj); return HB_ERR_UNSUPPORTED;                \
  }
  if (head == HB_HEAD_DISCRETE) { HB_HEAD_DISPATCH(discrete_head_kernel) }
  else { HB_HEAD_DISPATCH(box_head_kernel) }
#undef HB_HEAD_DISPATCH
  HB_LAUNCH_DONE(st, shape_label(MODE == MODE_GRAD ? "policy_head_grad" : MODE == MODE_EVAL ? "policy_head_eval" : "policy_head_act", a.rows, a.out, a.h));
  return HB_OK;
}

int launch_policy_head(int head, int mode, const HeadArgs& a, cudaStream_t st) {
  bool handled = false;
  const int rc = launch_policy_head_rows(head, mode, a, st, &handled);
  if (handled || rc != HB_OK) return rc;
  if (mode == MODE_ACT) return launch_head_mode<MODE_ACT>(head, a, st);
  if (mode == MODE_EVAL) return launch_head_mode<MODE_EVAL>(head, a, st);
  return launch_head_mode<MODE_GRAD>(head, a, st);
}

// ---- value head (v_net.py:65; v_critic.py:75-114)
__device__ __forceinline__ float huber_val(float e, float d, int use_huber, float* de) {
  if (!use_huber) { *de = e; return e * e / 2.f; }
  float ae = fabsf(e);
  if (ae <= d) { *de = e; return e * e / 2.f; }
  *de = e > 0.f ? d : -d;
  return d * (ae - d / 2.f);
}

template <int HPL, int GRAD>
__global__ void __launch_bounds__(ROW_THREADS) value_head_kernel(ValueArgs a) {
  __shared__ float sgw[256];
  __shared__ float sgb;
  __shared__ double sred[ROW_WARPS * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = a.h;
  if (GRAD) { for (int i = threadIdx.x; i < 256; i += ROW_THREADS) sgw[i] = 0.f; if (threadIdx.x == 0) sgb = 0.f; __syncthreads(); }
  float wv[HPL], gw[HPL];
#pragma unroll
  for (int q = 0; q < HPL; ++q) { int n = lane + 32 * q; wv[q] = n < h ? a.hw[n] : 0.f; gw[q] = 0.f; }
  const float bias = a.hbias[0];
  float vmean = 0.f, vstd = 1.f;
  if (GRAD && a.vn_state != nullptr) {  // valuenorm.py:38-45
    float d = fmaxf(a.vn_state[2], 1e-5f);
    float mu = a.vn_state[0] / d, msq = a.vn_state[1] / d;
    vmean = mu;
    vstd = sqrtf(fmaxf(msq - mu * mu, 1e-2f));
  }
  float gbias = 0.f;
  float lcg[HPL], lcb[HPL];
#pragma unroll
  for (int q = 0; q < HPL; ++q) lcg[q] = lcb[q] = 0.f;
  __shared__ float s_ln[GRAD ? 512 : 1];
  if (GRAD) { for (int i = threadIdx.x; i < 512; i += ROW_THREADS) s_ln[i] = 0.f; }
  double s_loss = 0.0, s_rows = 0.0;
  const int64_t w0 = (int64_t)blockIdx.x * ROW_WARPS + warp, nw = (int64_t)gridDim.x * ROW_WARPS;
  // software pipeline as in the policy heads: row r + nw's loads fly while row r is computed
  struct RowIn { float f[HPL], zr[HPL]; float mu, rs, vp, ret; };
  auto fetch = [&](int64_t r, int64_t src, RowIn& d) {
    load_feat<HPL>(a.feat + r * h, h, lane, d.f);
#pragma unroll
    for (int q = 0; q < HPL; ++q) d.zr[q] = 0.f;
    d.mu = d.rs = d.vp = d.ret = 0.f;
    if (GRAD) {
      if (a.ln_z != nullptr) { load_feat<HPL>(a.ln_z + r * h, h, lane, d.zr); d.mu = a.ln_stats[r * 2]; d.rs = a.ln_stats[r * 2 + 1]; }
      d.vp = a.value_preds[src];
      d.ret = a.returns[src];
    }
  };
  RowIn cur, nxt;
  int64_t src_cur = 0, src_nxt = 0;
  if (w0 < a.rows) { src_cur = (GRAD && a.index) ? (int64_t)a.index[w0] : w0; fetch(w0, src_cur, cur); }
  if (w0 + nw < a.rows) src_nxt = (GRAD && a.index) ? (int64_t)a.index[w0 + nw] : w0 + nw;
  for (int64_t r = w0; r < a.rows; r += nw) {
    int64_t src_nn = 0;
    if (r + 2 * nw < a.rows) src_nn = (GRAD && a.index) ? (int64_t)a.index[r + 2 * nw] : r + 2 * nw;
    if (r + nw < a.rows) fetch(r + nw, src_nxt, nxt);
    float f[HPL], zr[HPL];
#pragma unroll
    for (int q = 0; q < HPL; ++q) { f[q] = cur.f[q]; zr[q] = cur.zr[q]; }
    const float ln_mu = cur.mu, ln_rs = cur.rs, vp = cur.vp;
    float ret = cur.ret;
    float p = 0.f;
#pragma unroll
    for (int q = 0; q < HPL; ++q) p = fmaf(f[q], wv[q], p);
    const float v = warp_sum(p) + bias;
    if (!GRAD) { if (lane == 0) a.values_out[r] = v; cur = nxt; src_cur = src_nxt; src_nxt = src_nn; continue; }
    if (a.vn_state != nullptr) ret = (ret - vmean) / vstd;
    const float dv = v - vp;
    const float dc = fminf(fmaxf(dv, -a.clip), a.clip);
    const float vclip = vp + dc;
    const bool pass = dv >= -a.clip && dv <= a.clip;
    float de_c, de_o;
    const float l_c = huber_val(ret - vclip, a.huber_delta, a.use_huber, &de_c);
    const float l_o = huber_val(ret - v, a.huber_delta, a.use_huber, &de_o);
    float g_o = -de_o, g_c = pass ? -de_c : 0.f;  // d l / d v
    float loss = l_o, g = g_o;
    if (a.use_clipped) {
      if (l_c > l_o) { loss = l_c; g = g_c; }
      else if (l_c == l_o) { loss = l_o; g = 0.5f * (g_o + g_c); }
    }
    g *= a.coef;
    if (lane == 0) { s_loss += (double)loss; s_rows += 1.0; }
    gbias += g;
    float df[HPL];
#pragma unroll
    for (int q = 0; q < HPL; ++q) {
      int n = lane + 32 * q;
      df[q] = 0.f;
      if (n < h) { df[q] = g * wv[q]; gw[q] = fmaf(g, f[q], gw[q]); if (a.ln_z == nullptr) a.dfeat[r * h + n] = df[q]; }
    }
    if (a.ln_z != nullptr)
      ln_act_bwd_row<HPL>(df, zr, ln_mu, ln_rs, a.ln_w, h, a.ln_act, lane, a.dfeat + r * h, lcg, lcb);
    cur = nxt; src_cur = src_nxt; src_nxt = src_nn;
  }
  if (GRAD) {
#pragma unroll
    for (int q = 0; q < HPL; ++q) { int n = lane + 32 * q; if (n < h) atomicAdd(&sgw[n], gw[q]); }
    if (lane == 0) atomicAdd(&sgb, gbias);
    __syncthreads();
    const int64_t slot = a.part_stride ? a.part_delta + (int64_t)blockIdx.x * a.part_stride : 0;
    for (int n = threadIdx.x; n < h; n += ROW_THREADS) acc_out(a.g_hw + n, sgw[n], slot);
    if (threadIdx.x == 0) acc_out(a.g_hbias, sgb, slot);
    block_add_scalars(s_loss, s_rows, 0.0, 0.0, a.scalars, sred);
    if (a.ln_z != nullptr) ln_affine_flush<HPL>(lcg, lcb, h, lane, s_ln, a.g_ln_w, a.g_ln_b, slot);
  }
}

int launch_value_head(int grad, const ValueArgs& a, cudaStream_t st) {
  if (a.rows <= 0) return HB_OK;
  bool handled = false;
  const int rc = launch_value_head_rows(grad, a, st, &handled);
  if (handled || rc != HB_OK) return rc;
  const int hpl = a.h <= 32 ? 1 : a.h <= 64 ? 2 : a.h <= 128 ? 4 : 8;
  const int g = (grad && a.part_stride) ? slot_grid(a.rows) : row_grid(a.rows);
#define HB_V(H)                                                                 \
  case H:                                                                       \
    if (grad) value_head_kernel<H, 1><<<g, ROW_THREADS, 0, st>>>(a);            \
    else value_head_kernel<H, 0><<<g, ROW_THREADS, 0, st>>>(a);                 \
    break;
  switch (hpl) { HB_V(1) HB_V(2) HB_V(4) HB_V(8) }
#undef HB_V
  HB_LAUNCH_DONE(st, shape_label(grad ? "value_head_grad" : "value_head_fwd", a.rows, 1, a.h));
  return HB_OK;
}

}  // namespace hb
