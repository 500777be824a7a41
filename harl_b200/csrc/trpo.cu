// Trust-region (HATRPO) kernels: the Fisher-vector product in Gauss-Newton form, the KL / surrogate evaluation of
// the backtracking line search, and the conjugate-gradient vector algebra (sm_100a).
//
// Reference: harl/algorithms/actors/hatrpo.py:37-194 and harl/utils/trpo_util.py:49-158.  The reference forms the
// Fisher-vector product by differentiating mean KL(pi || pi) twice.  At new == old the first derivative of the KL
// w.r.t. the distribution parameters is zero, so that Hessian is exactly J^T H J with
//   Categorical (kl_approx over ALL normalised logits, masked ones included): H = I
//   DiagGaussian: H = diag(1 / sigma^2) over the means, 2 over log sigma
// and one product is: a forward-mode (tangent) pass -> H / rows -> the ordinary backward pass.
//
//   tangent_prepare      : tangent of the derived weights (W^T, feature-norm fold) for a parameter-space vector
//   jvp_linear_ln        : tangent of one Linear -> act -> LayerNorm block (two GEMMs into one tile + LN tangent)
//   trpo_head (3 modes)  : OLD  stores the old distribution (normalised logits / means),
//                          FVP  head tangent -> H/rows -> head backward (+ fused LN/act backward of the last block),
//                          LS   surrogate, entropy, ratio and KL(old || new) sums of one line-search trial
//   cg_* / full_step / apply_step / fvp_finish : single-CTA vector kernels (<= ~100k parameters)
#include <math.h>

#include "common.cuh"
#include "gemm_tile.cuh"
#include "kernels.cuh"
#include "row_helpers.cuh"
#include "trpo.cuh"

namespace hb {

// ------------------------------------------------------------------ tangent of the prepared weights
__global__ void tangent_prepare_kernel(ParamLayout P, PrepLayout Q, int feature_norm, int head, int out_dim,
                                       const float* __restrict__ params, const float* __restrict__ v,
                                       float* __restrict__ tprep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q.tk[0]) return;
  for (int l = 0; l < Q.n_layers; ++l) {
    int n = Q.n[l], K = Q.k[l], kp = Q.kpad[l];
    if (i >= Q.wt[l] && i < Q.wt[l] + kp * n) {
      int k = (i - Q.wt[l]) / n, j = (i - Q.wt[l]) % n;
      float t = 0.f;
      if (k < K) {
        t = v[P.w[l] + j * K + k];
        if (l == 0 && feature_norm) t = t * params[P.fn_w + k] + params[P.w[0] + j * K + k] * v[P.fn_w + k];
      }
      tprep[i] = t;
      return;
    }
    if (i >= Q.bias[l] && i < Q.bias[l] + n) {
      int j = i - Q.bias[l];
      float t = v[P.b[l] + j];
      if (l == 0 && feature_norm)
        for (int k = 0; k < K; ++k)
          t += v[P.w[0] + j * K + k] * params[P.fn_b + k] + params[P.w[0] + j * K + k] * v[P.fn_b + k];
      tprep[i] = t;
      return;
    }
    if (i >= Q.lnw[l] && i < Q.lnw[l] + n) { tprep[i] = v[P.lnw[l] + i - Q.lnw[l]]; return; }
    if (i >= Q.lnb[l] && i < Q.lnb[l] + n) { tprep[i] = v[P.lnb[l] + i - Q.lnb[l]]; return; }
  }
  int h = Q.n[Q.n_layers - 1];
  if (i >= Q.hw && i < Q.hw + out_dim * h) { tprep[i] = v[P.hw + i - Q.hw]; return; }
  if (i >= Q.hbias && i < Q.hbias + out_dim) { tprep[i] = v[P.hbias + i - Q.hbias]; return; }
  if (head == HB_HEAD_BOX && i >= Q.log_std && i < Q.log_std + out_dim) { tprep[i] = v[P.log_std + i - Q.log_std]; return; }
  for (int r = 0; r < Q.rnn_layers; ++r) {  // linear in the parameters: the same transposes as hb_net_prepare
    const int g3 = 3 * h;
    if (i >= Q.rnn_wih_t[r] && i < Q.rnn_wih_t[r] + h * g3) {
      int k = (i - Q.rnn_wih_t[r]) / g3, j = (i - Q.rnn_wih_t[r]) % g3;
      tprep[i] = v[P.rnn_wih[r] + j * h + k];
      return;
    }
    if (i >= Q.rnn_whh_t[r] && i < Q.rnn_whh_t[r] + h * g3) {
      int k = (i - Q.rnn_whh_t[r]) / g3, j = (i - Q.rnn_whh_t[r]) % g3;
      tprep[i] = v[P.rnn_whh[r] + j * h + k];
      return;
    }
    if (i >= Q.rnn_bih[r] && i < Q.rnn_bih[r] + g3) { tprep[i] = v[P.rnn_bih[r] + i - Q.rnn_bih[r]]; return; }
    if (i >= Q.rnn_bhh[r] && i < Q.rnn_bhh[r] + g3) { tprep[i] = v[P.rnn_bhh[r] + i - Q.rnn_bhh[r]]; return; }
  }
  if (Q.rnn_layers) {
    if (i >= Q.rnn_lnw && i < Q.rnn_lnw + h) { tprep[i] = v[P.rnn_lnw + i - Q.rnn_lnw]; return; }
    if (i >= Q.rnn_lnb && i < Q.rnn_lnb + h) { tprep[i] = v[P.rnn_lnb + i - Q.rnn_lnb]; return; }
  }
  tprep[i] = 0.f;
}

int launch_tangent_prepare(const hb_net_desc* d, const ParamLayout& P, const PrepLayout& Q, const float* params,
                           const float* v, float* tprep, cudaStream_t st) {
  tangent_prepare_kernel<<<(Q.tk[0] + 255) / 256, 256, 0, st>>>(P, Q, d->feature_norm, d->head, d->out_dim, params, v, tprep);
  HB_LAUNCH_DONE(st, "trpo_tangent_prepare");
  return HB_OK;
}

// ------------------------------------------------------------------ tangent of one trunk block
// zd = Xd WT + X WdT + bd;  ad = act'(Z) zd;  xh = (act(Z) - mu) rstd;
// yd = gd xh + betad + gamma rstd (ad - mean(ad) - xh mean(xh ad))
template <int NT, int ACT>
__global__ void __launch_bounds__(256) jvp_linear_ln_kernel(const float* __restrict__ X, int ldx,
                                                            const float* __restrict__ Xd,
                                                            const float* __restrict__ WT, const float* __restrict__ WdT,
                                                            const float* __restrict__ bd, const float* __restrict__ lnw,
                                                            const float* __restrict__ lnwd, const float* __restrict__ lnbd,
                                                            const float* __restrict__ Z, const float* __restrict__ stats,
                                                            float* __restrict__ Yd, int64_t M, int N, int Kred) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmSmem<NT>& s = *reinterpret_cast<GemmSmem<NT>*>(smem_raw);
  constexpr int NCH = NT / 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  float acc[4][NT / 16];
  gemm_mainloop<NT, true>(X, ldx, WdT, N, M, Kred, N, row0, s, acc);
  if (Xd != nullptr) gemm_mainloop<NT, false>(Xd, ldx, WT, N, M, Kred, N, row0, s, acc);

  float4 bv[NCH], gw[NCH], gwd[NCH], gbd[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    int n = c * 64 + tx * 4;
    bv[c] = gw[c] = gwd[c] = gbd[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
      bv[c] = *reinterpret_cast<const float4*>(bd + n);
      gw[c] = *reinterpret_cast<const float4*>(lnw + n);
      gwd[c] = *reinterpret_cast<const float4*>(lnwd + n);
      gbd[c] = *reinterpret_cast<const float4*>(lnbd + n);
    }
  }
  const float inv_n = 1.f / (float)N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    const bool rok = row < M;
    float mu = 0.f, rstd = 0.f;
    if (rok) { mu = stats[row * 2]; rstd = stats[row * 2 + 1]; }
    float xh[NCH * 4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int n = c * 64 + tx * 4;
      const bool ok = rok && n < N;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) z = *reinterpret_cast<const float4*>(Z + row * N + n);
      const float zz[4] = {z.x, z.y, z.z, z.w};
      const float bb[4] = {bv[c].x, bv[c].y, bv[c].z, bv[c].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float ad = 0.f, x = 0.f;
        if (ok) {
          ad = act_bwd<ACT>(zz[j]) * (acc[i][c * 4 + j] + bb[j]);
          x = (act_fwd<ACT>(zz[j]) - mu) * rstd;
        }
        acc[i][c * 4 + j] = ad;
        xh[c * 4 + j] = x;
        s1 += ad;
        s2 = fmaf(ad, x, s2);
      }
    }
    const float m1 = half_warp_sum(s1) * inv_n, m2 = half_warp_sum(s2) * inv_n;
    if (rok) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        int n = c * 64 + tx * 4;
        if (n < N) {
          float4 o;
          o.x = gwd[c].x * xh[c * 4 + 0] + gbd[c].x + gw[c].x * rstd * (acc[i][c * 4 + 0] - m1 - xh[c * 4 + 0] * m2);
          o.y = gwd[c].y * xh[c * 4 + 1] + gbd[c].y + gw[c].y * rstd * (acc[i][c * 4 + 1] - m1 - xh[c * 4 + 1] * m2);
          o.z = gwd[c].z * xh[c * 4 + 2] + gbd[c].z + gw[c].z * rstd * (acc[i][c * 4 + 2] - m1 - xh[c * 4 + 2] * m2);
          o.w = gwd[c].w * xh[c * 4 + 3] + gbd[c].w + gw[c].w * rstd * (acc[i][c * 4 + 3] - m1 - xh[c * 4 + 3] * m2);
          *reinterpret_cast<float4*>(Yd + row * N + n) = o;
        }
      }
    }
  }
}

template <int NT>
static int launch_jvp_nt(int act, const float* X, int ldx, const float* Xd, const float* WT, const float* WdT,
                         const float* bd, const float* lnw, const float* lnwd, const float* lnbd, const float* Z,
                         const float* stats, float* Yd, int64_t M, int N, int Kred, cudaStream_t st) {
  const size_t smem = sizeof(GemmSmem<NT>);
  dim3 grid((unsigned)ceil_div64(M, BM));
#define HB_JVP_CASE(A)                                                                                          \
  case A: {                                                                                                     \
    auto kern = jvp_linear_ln_kernel<NT, A>;                                                                    \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    kern<<<grid, 256, smem, st>>>(X, ldx, Xd, WT, WdT, bd, lnw, lnwd, lnbd, Z, stats, Yd, M, N, Kred);          \
  } break;
  switch (act) {
    HB_JVP_CASE(HB_ACT_RELU) HB_JVP_CASE(HB_ACT_TANH) HB_JVP_CASE(HB_ACT_SIGMOID) HB_JVP_CASE(HB_ACT_LEAKY_RELU)
    HB_JVP_CASE(HB_ACT_SELU) HB_JVP_CASE(HB_ACT_HARDSWISH) HB_JVP_CASE(HB_ACT_IDENTITY)
    default: set_error("activation %d", act); return HB_ERR_UNSUPPORTED;
  }
#undef HB_JVP_CASE
  HB_LAUNCH_DONE(st, shape_label("trpo_jvp_linear_ln", M, N, Kred));
  return HB_OK;
}

int launch_jvp_linear_ln(int act, const float* X, int ldx, const float* Xd, const float* WT, const float* WdT,
                         const float* bd, const float* lnw, const float* lnwd, const float* lnbd, const float* Z,
                         const float* stats, float* Yd, int64_t M, int N, int Kred, cudaStream_t st) {
  if (M <= 0) return HB_OK;
  if (N <= 64) return launch_jvp_nt<64>(act, X, ldx, Xd, WT, WdT, bd, lnw, lnwd, lnbd, Z, stats, Yd, M, N, Kred, st);
  if (N <= 128) return launch_jvp_nt<128>(act, X, ldx, Xd, WT, WdT, bd, lnw, lnwd, lnbd, Z, stats, Yd, M, N, Kred, st);
  return launch_jvp_nt<256>(act, X, ldx, Xd, WT, WdT, bd, lnw, lnwd, lnbd, Z, stats, Yd, M, N, Kred, st);
}

// ------------------------------------------------------------------ head kernel (one warp per row)
// Up to 8 outputs: out[j] = fa . Wa[j] (+ fb . Wb[j]) + bias[j] lands in lane j.  The 8 per-lane partial sums are
// combined by a reduce-scatter (4 + 2 + 1 exchanges while the live set halves, then 2 plain butterfly steps and one
// gather): 10 shuffles instead of the 40 of eight independent butterflies -- ncu showed the first cut of these kernels
// issue-bound at ~1100 warp instructions per row, most of them shuffle/add pairs of the head reductions.
template <int HPL, bool TWO>
__device__ __forceinline__ float head_linear_rs8(const float (&fa)[HPL], const float* __restrict__ swa,
                                                 const float (&fb)[HPL], const float* __restrict__ swb, int h, int nout,
                                                 const float* __restrict__ sbias, int lane) {
  float p[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    p[j] = 0.f;
    if (j < nout) {
#pragma unroll
      for (int q = 0; q < HPL; ++q) {
        const int n = lane + 32 * q;
        if (n < h) {
          p[j] = fmaf(fa[q], swa[j * h + n], p[j]);
          if (TWO) p[j] = fmaf(fb[q], swb[j * h + n], p[j]);
        }
      }
    }
  }
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
  float q4[4], q2[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b16 ? p[i] : p[i + 4], keep = b16 ? p[i + 4] : p[i];
    q4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b8 ? q4[i] : q4[i + 2], keep = b8 ? q4[i + 2] : q4[i];
    q2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float q1 = (b4 ? q2[1] : q2[0]) + __shfl_xor_sync(0xffffffffu, b4 ? q2[0] : q2[1], 4);
  q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
  q1 += __shfl_xor_sync(0xffffffffu, q1, 1);   // lanes 4j .. 4j+3 hold output j
  const float v = __shfl_sync(0xffffffffu, q1, (lane & 7) << 2);
  return lane < nout ? v + sbias[lane] : 0.f;
}

// lane j < out holds output j.  Shared memory: hw [out][h], hwd [out][h] (FVP), bias[32], biasd[32], zero[32],
// grad accumulators [out][h] + [32] (FVP), double scratch.
template <int HPL, int MAXJ, int HEAD, int MODE>
__global__ void __launch_bounds__(ROW_THREADS, (HPL * MAXJ <= 32) ? 2 : 1) trpo_head_kernel(TrpoHeadArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int h = a.h, na = a.out;
  float* shw = sm;                                        // [out][h]
  float* shwd = shw + na * h;                             // [out][h]  (FVP)
  float* sb = shwd + (MODE == TR_FVP ? na * h : 0);       // [32]
  float* sbd = sb + 32;                                   // [32]
  float* szero = sbd + 32;                                // [32]
  float* sg = szero + 32;                                 // [out][h]  (FVP)
  float* sgb = sg + (MODE == TR_FVP ? na * h : 0);        // [32]
  double* sred = reinterpret_cast<double*>(sgb + 32);     // [ROW_WARPS * 4]
  __shared__ float s_ln[MODE == TR_FVP ? 512 : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < na * h; i += ROW_THREADS) {
    shw[i] = a.hw[i];
    if (MODE == TR_FVP) { shwd[i] = a.hwd[i]; sg[i] = 0.f; }
  }
  if (threadIdx.x < 32) {
    sb[threadIdx.x] = threadIdx.x < na ? a.hbias[threadIdx.x] : 0.f;
    sbd[threadIdx.x] = (MODE == TR_FVP && threadIdx.x < na) ? a.hbd[threadIdx.x] : 0.f;
    szero[threadIdx.x] = 0.f;
    sgb[threadIdx.x] = 0.f;
  }
  if (MODE == TR_FVP) { for (int i = threadIdx.x; i < 512; i += ROW_THREADS) s_ln[i] = 0.f; }
  __syncthreads();
  const bool valid = lane < na;
  // DiagGaussian scale of the CURRENT parameters (distributions.py:86-89)
  float std = 1.f, log_std_v = 0.f, std_old = 1.f;
  if (HEAD == HB_HEAD_BOX && valid) {
    std = a.std_y / (1.f + expf(-a.log_std[lane] / a.std_x));
    log_std_v = logf(std);
    if (MODE == TR_LS) std_old = a.std_y / (1.f + expf(-a.old_log_std[lane] / a.std_x));
  }
  const float ent_box = HEAD == HB_HEAD_BOX ? warp_sum(valid ? 0.5f + 0.5f * HB_LOG_2PI_F + log_std_v : 0.f) : 0.f;
  float gacc[MODE == TR_FVP ? MAXJ : 1][HPL];
  float gb = 0.f;
  float lcg[HPL], lcb[HPL];
#pragma unroll
  for (int q = 0; q < HPL; ++q) lcg[q] = lcb[q] = 0.f;
  if constexpr (MODE == TR_FVP) {
#pragma unroll
    for (int j = 0; j < MAXJ; ++j)
#pragma unroll
      for (int q = 0; q < HPL; ++q) gacc[j][q] = 0.f;
  }
  double s_loss = 0.0, s_ent = 0.0, s_ratio = 0.0, s_kl = 0.0;
  const int64_t w0 = (int64_t)blockIdx.x * ROW_WARPS + warp, nw = (int64_t)gridDim.x * ROW_WARPS;
  // Software pipeline (as in the policy head kernels): the loads of row r + nw are in flight while row r is computed,
  // so a warp keeps two rows' worth of memory requests outstanding instead of one.
  struct RowIn { float f[HPL], fd[MODE == TR_FVP ? HPL : 1], zr[MODE == TR_FVP ? HPL : 1]; float mu, rs, od, av; int64_t src; };
  auto fetch = [&](int64_t r, RowIn& d) {
    d.src = a.index ? (int64_t)a.index[r] : r;
    load_feat<HPL>(a.feat + r * h, h, lane, d.f);
    d.mu = d.rs = d.od = 0.f;
    d.av = 1.f;
    if (HEAD == HB_HEAD_DISCRETE && valid && a.avail != nullptr) d.av = a.avail[d.src * na + lane];
    if constexpr (MODE == TR_FVP) {
      load_feat<HPL>(a.featd + r * h, h, lane, d.fd);
      load_feat<HPL>(a.ln_z + r * h, h, lane, d.zr);
      d.mu = a.ln_stats[r * 2];
      d.rs = a.ln_stats[r * 2 + 1];
    }
    if (MODE != TR_OLD && valid) d.od = a.old_dist[r * na + lane];
  };
  RowIn cur, nxt;
  if (w0 < a.rows) fetch(w0, cur);
  for (int64_t r = w0; r < a.rows; r += nw) {
    if (r + nw < a.rows) fetch(r + nw, nxt);
    const int64_t src = cur.src;
    float f[HPL];
#pragma unroll
    for (int q = 0; q < HPL; ++q) f[q] = cur.f[q];
    const bool masked = HEAD == HB_HEAD_DISCRETE && valid && cur.av == 0.f;
    const float od_lane = cur.od;
    (void)src; (void)od_lane;
    if constexpr (MODE == TR_OLD) {
      float o = MAXJ <= 8 ? head_linear_rs8<HPL, false>(f, shw, f, shw, h, na, sb, lane)
                          : head_linear<HPL, 8>(f, shw, h, na, sb, lane);
      if (HEAD == HB_HEAD_DISCRETE) {
        if (masked) o = -1e10f;
        const float mx = warp_max(valid ? o : -INFINITY);
        const float lse = mx + logf(warp_sum(valid ? expf(o - mx) : 0.f));
        o = o - lse;
      }
      if (valid) a.old_dist_out[r * na + lane] = o;
    } else if constexpr (MODE == TR_FVP) {
      float fd[HPL], zr[HPL];
#pragma unroll
      for (int q = 0; q < HPL; ++q) { fd[q] = cur.fd[q]; zr[q] = cur.zr[q]; }
      const float ln_mu = cur.mu, ln_rs = cur.rs;
      float zd = MAXJ <= 8 ? head_linear_rs8<HPL, true>(fd, shw, f, shwd, h, na, sbd, lane)
                           : head_linear<HPL, 8>(fd, shw, h, na, szero, lane) + head_linear<HPL, 8>(f, shwd, h, na, sbd, lane);
      float dl = 0.f;
      if (HEAD == HB_HEAD_DISCRETE) {
        if (masked || !valid) zd = 0.f;                       // a masked logit is the constant -1e10
        const float p = valid ? expf(od_lane) : 0.f;
        const float sdot = warp_sum(p * zd);                  // d logsumexp
        const float u = valid ? (zd - sdot) * a.inv_rows : 0.f;  // H = I over every normalised logit
        const float usum = warp_sum(u);
        if (valid && !masked) dl = u - p * usum;              // back through log_softmax to the raw logits
      } else {
        if (valid) dl = zd / (std * std) * a.inv_rows;        // H = 1 / sigma^2 over the means
      }
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
      ln_act_bwd_row<HPL>(df, zr, ln_mu, ln_rs, a.ln_w, h, a.ln_act, lane, a.dfeat + r * h, lcg, lcb);
    } else {
      // ---- TR_LS: hatrpo.py:142-181 for one candidate step
      const float w = a.use_active ? a.active[src] : 1.f;
      const float fac = a.factor ? a.factor[src] : 1.f;
      const float adv = a.adv[src];
      float o = MAXJ <= 8 ? head_linear_rs8<HPL, false>(f, shw, f, shw, h, na, sb, lane)
                          : head_linear<HPL, 8>(f, shw, h, na, sb, lane);
      float ratio, ent;
      double klrow;
      if (HEAD == HB_HEAD_DISCRETE) {
        if (masked) o = -1e10f;
        const float mx = warp_max(valid ? o : -INFINITY);
        const float lse = mx + logf(warp_sum(valid ? expf(o - mx) : 0.f));
        const float lq = valid ? o - lse : 0.f;
        const float lp = valid ? od_lane : 0.f;
        const float klj = valid ? (expf(lq - lp) - 1.f - lq) + lp : 0.f;  // kl_approx(p_old, q_new), trpo_util.py:49-53
        klrow = (double)warp_sum(klj);
        const int act = (int)a.actions[src];
        const float lpa = __shfl_sync(0xffffffffu, lq, act);
        ratio = expf(lpa - a.old_logp[src]);
        ent = -warp_sum(valid ? fmaxf(lq, -3.4028234663852886e38f) * expf(lq) : 0.f);
      } else {
        const float act = valid ? a.actions[src * na + lane] : 0.f;
        const float dlt = act - o;
        const float lp = -(dlt * dlt) / (2.f * std * std) - log_std_v - 0.5f * HB_LOG_2PI_F;
        const float e = valid ? expf(lp - a.old_logp[src * na + lane]) : (a.agg_prod ? 1.f : 0.f);
        if (a.agg_prod) {
          ratio = e;
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) ratio *= __shfl_xor_sync(0xffffffffu, ratio, s);
        } else {
          ratio = warp_sum(e) / (float)na;
        }
        double klj = 0.0;
        if (valid) {  // _kl_normal_normal in float64, trpo_util.py:56-62 (p = old, q = new)
          const double sp = (double)std_old, sq = (double)std;
          const double vr = (sp / sq) * (sp / sq);
          const double t1 = (((double)od_lane - (double)o) / sq);
          klj = 0.5 * (vr + t1 * t1 - 1.0 - log(vr));
        }
        klrow = warp_sum_d(klj);
        ent = ent_box;
      }
      if (lane == 0) {
        s_loss += (double)(ratio * fac * adv * w);
        s_ent += (double)(ent * w);
        s_ratio += (double)ratio;
        s_kl += klrow;
      }
    }
    cur = nxt;
  }
  if constexpr (MODE == TR_FVP) {
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
    ln_affine_flush<HPL>(lcg, lcb, h, lane, s_ln, a.g_ln_w, a.g_ln_b, slot);
  }
  if constexpr (MODE == TR_LS) block_add_scalars(s_loss, s_ent, s_ratio, s_kl, a.scalars, sred);
}

template <int HEAD, int MODE>
static int launch_trpo_head_mode(const TrpoHeadArgs& a, cudaStream_t st) {
  const int hpl = a.h <= 32 ? 1 : a.h <= 64 ? 2 : a.h <= 128 ? 4 : 8;
  // MAXJ selects the head reduction in every mode (<= 8 outputs: reduce-scatter; wider heads: the generic butterflies)
  // and sizes the per-lane gradient accumulators of the FVP mode
  const int maxj = a.out <= 8 ? 8 : a.out <= 16 ? 16 : 32;
  if (MODE == TR_FVP) {
    if (maxj * hpl > 64) {
      set_error("trust-region head: out_dim %d with hidden %d exceeds the register-accumulator budget", a.out, a.h);
      return HB_ERR_UNSUPPORTED;
    }
  }
  const size_t smem = (size_t)(a.out * a.h) * 4 * (MODE == TR_FVP ? 3 : 1) + 5 * 32 * 4 + ROW_WARPS * 4 * sizeof(double) + 16;
  const int grid = (MODE == TR_FVP && a.part_stride) ? slot_grid(a.rows) : row_grid(a.rows);
#define HB_TR_LAUNCH(HPLV, MAXJV)                                                                              \
  do {                                                                                                         \
    auto kern = trpo_head_kernel<HPLV, MAXJV, HEAD, MODE>;                                                     \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    kern<<<grid, ROW_THREADS, smem, st>>>(a);                                                                  \
  } while (0)
  switch (hpl * 100 + maxj) {
    case 108: HB_TR_LAUNCH(1, 8); break;
    case 116: HB_TR_LAUNCH(1, 16); break;
    case 132: HB_TR_LAUNCH(1, 32); break;
    case 208: HB_TR_LAUNCH(2, 8); break;
    case 216: HB_TR_LAUNCH(2, 16); break;
    case 232: HB_TR_LAUNCH(2, 32); break;
    case 408: HB_TR_LAUNCH(4, 8); break;
    case 416: HB_TR_LAUNCH(4, 16); break;
    case 808: HB_TR_LAUNCH(8, 8); break;
    default: set_error("trust-region head dispatch %d/%d", hpl, maxj); return HB_ERR_UNSUPPORTED;
  }
#undef HB_TR_LAUNCH
  HB_LAUNCH_DONE(st, shape_label(MODE == TR_FVP ? "trpo_head_fvp" : MODE == TR_LS ? "trpo_head_linesearch" : "trpo_head_old_dist",
                                 a.rows, a.out, a.h));
  return HB_OK;
}

int launch_trpo_head(int head, int mode, const TrpoHeadArgs& a, cudaStream_t st) {
  if (a.rows <= 0) return HB_OK;
  if (head == HB_HEAD_DISCRETE) {
    if (mode == TR_OLD) return launch_trpo_head_mode<HB_HEAD_DISCRETE, TR_OLD>(a, st);
    if (mode == TR_FVP) return launch_trpo_head_mode<HB_HEAD_DISCRETE, TR_FVP>(a, st);
    return launch_trpo_head_mode<HB_HEAD_DISCRETE, TR_LS>(a, st);
  }
  if (head == HB_HEAD_BOX) {
    if (mode == TR_OLD) return launch_trpo_head_mode<HB_HEAD_BOX, TR_OLD>(a, st);
    if (mode == TR_FVP) return launch_trpo_head_mode<HB_HEAD_BOX, TR_FVP>(a, st);
    return launch_trpo_head_mode<HB_HEAD_BOX, TR_LS>(a, st);
  }
  set_error("trust-region update needs a policy head");
  return HB_ERR_INVALID;
}

// ------------------------------------------------------------------ vector kernels (one CTA of 1024 threads)
__device__ __forceinline__ double block_sum_1024(double v, double* sred) {
  v = warp_sum_d(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // sred may still be read from a previous reduction
  if (lane == 0) sred[warp] = v;
  __syncthreads();
  double t = threadIdx.x < 32 ? sred[threadIdx.x] : 0.0;
  if (warp == 0) {
    t = warp_sum_d(t);
    if (lane == 0) sred[32] = t;
  }
  __syncthreads();
  return sred[32];
}

// trpo_util.py:116-119.  state = {rdotr, done}
__global__ void __launch_bounds__(1024) cg_init_kernel(const float* __restrict__ b, float* x, float* r, float* p,
                                                       float* state, int n) {
  __shared__ double sred[33];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) {
    const float bi = b[i];
    x[i] = 0.f; r[i] = bi; p[i] = bi;
    acc += (double)bi * (double)bi;
  }
  const double rr = block_sum_1024(acc, sred);
  if (threadIdx.x == 0) { state[0] = (float)rr; state[1] = 0.f; }
}

// trpo_util.py:120-132 (one iteration; a no-op once rdotr fell below residual_tol)
__global__ void __launch_bounds__(1024) cg_step_kernel(float* p, const float* __restrict__ avp, float* x, float* r,
                                                       float* state, int n, float tol) {
  __shared__ double sred[33];
  if (state[1] != 0.f) return;
  const float rdotr = state[0];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += (double)p[i] * (double)avp[i];
  const float pap = (float)block_sum_1024(acc, sred);
  const float alpha = rdotr / pap;
  acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) {
    x[i] = fmaf(alpha, p[i], x[i]);
    const float ri = r[i] - alpha * avp[i];
    r[i] = ri;
    acc += (double)ri * (double)ri;
  }
  const float new_rdotr = (float)block_sum_1024(acc, sred);
  const float beta = new_rdotr / rdotr;
  for (int i = threadIdx.x; i < n; i += 1024) p[i] = fmaf(beta, p[i], r[i]);
  if (threadIdx.x == 0) { state[0] = new_rdotr; if (new_rdotr < tol) state[1] = 1.f; }
}

// hatrpo.py:123-133: shs = 0.5 x.Fx; step_size = 1/sqrt(shs/kl_threshold); full = step_size x; expected = g.full
__global__ void __launch_bounds__(1024) full_step_kernel(const float* __restrict__ x, const float* __restrict__ fx,
                                                         const float* __restrict__ g, float kl_threshold,
                                                         float* __restrict__ full, double* out3, int n) {
  __shared__ double sred[33];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += (double)x[i] * (double)fx[i];
  const float shs = 0.5f * (float)block_sum_1024(acc, sred);
  const float step_size = 1.f / sqrtf(shs / kl_threshold);
  acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) {
    const float fi = step_size * x[i];
    full[i] = fi;
    acc += (double)g[i] * (double)fi;
  }
  const double expected = block_sum_1024(acc, sred);
  if (threadIdx.x == 0) { out3[0] = (double)shs; out3[1] = (double)step_size; out3[2] = (double)(float)expected; }
}

__global__ void apply_step_kernel(float* __restrict__ params, const float* __restrict__ params0,
                                  const float* __restrict__ full, float fraction, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) params[i] = params0[i] + fraction * full[i];
}

__global__ void vec_scale_kernel(float* __restrict__ x, float s, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s;
}

// out += damping * v, and for a DiagGaussian head the log_std block: 2 (d log sigma / d log_std)^2 v (the KL is the
// same for every row, so the row mean is the per-row value)
__global__ void fvp_finish_kernel(float* __restrict__ out, const float* __restrict__ v, const float* __restrict__ params,
                                  float damping, int n, int ls_off, int ls_n, float std_x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o = out[i] + damping * v[i];
  if (i >= ls_off && i < ls_off + ls_n) {
    const float sig = 1.f / (1.f + expf(-params[i] / std_x));
    const float ds = (1.f - sig) / std_x;  // d log(sigmoid(p / x) y) / dp
    o += 2.f * ds * ds * v[i];
  }
  out[i] = o;
}

int launch_cg_init(const float* b, float* x, float* r, float* p, float* state, int n, cudaStream_t st) {
  cg_init_kernel<<<1, 1024, 0, st>>>(b, x, r, p, state, n);
  HB_LAUNCH_DONE(st, "trpo_cg_init");
  return HB_OK;
}
int launch_cg_step(float* p, const float* avp, float* x, float* r, float* state, int n, float tol, cudaStream_t st) {
  cg_step_kernel<<<1, 1024, 0, st>>>(p, avp, x, r, state, n, tol);
  HB_LAUNCH_DONE(st, "trpo_cg_step");
  return HB_OK;
}
int launch_full_step(const float* x, const float* fx, const float* g, float kl_threshold, float* full, double* out3, int n,
                     cudaStream_t st) {
  full_step_kernel<<<1, 1024, 0, st>>>(x, fx, g, kl_threshold, full, out3, n);
  HB_LAUNCH_DONE(st, "trpo_full_step");
  return HB_OK;
}
int launch_apply_step(float* params, const float* params0, const float* full, float fraction, int n, cudaStream_t st) {
  apply_step_kernel<<<(n + 255) / 256, 256, 0, st>>>(params, params0, full, fraction, n);
  HB_LAUNCH_DONE(st, "trpo_apply_step");
  return HB_OK;
}
int launch_vec_scale(float* x, float s, int n, cudaStream_t st) {
  vec_scale_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, s, n);
  HB_LAUNCH_DONE(st, "vec_scale");
  return HB_OK;
}
int launch_fvp_finish(float* out, const float* v, const float* params, float damping, int n, int ls_off, int ls_n,
                      float std_x, cudaStream_t st) {
  fvp_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, v, params, damping, n, ls_off, ls_n, std_x);
  HB_LAUNCH_DONE(st, "trpo_fvp_finish");
  return HB_OK;
}

}  // namespace hb
