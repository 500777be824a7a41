// FP32 SIMT GEMM kernels with fused LayerNorm / activation epilogues (sm_100a).
//
//   linear_ln_fwd : Y = LN(act(X WT + b))            mlp.py:25-36 (one MLPLayer block)
//   dx_ln_bwd     : dZp = act'(Zp) * LNbwd(dZ W)      autograd of the block above (happo.py:91)
//   dw_accum      : dW += dZ^T X, db += colsum(dZ)
//
// Tiling: CTA = 256 threads = 16 (tx) x 16 (ty); tile = 64 rows x NT cols, NT in
// {64,128,256} covering the whole output row so LayerNorm statistics stay inside one
// half-warp (16 lanes).  Thread (tx,ty) owns rows ty*4..+3 and the float4 column chunks
// c*64 + tx*4 (c < NT/64): consecutive lanes read consecutive float4 of the B tile
// (conflict-free LDS.128) and broadcast the A operand.  Reduction tiles of 16, register
// prefetch + two shared buffers, one __syncthreads per tile.
#include "common.cuh"
#include "gemm_tile.cuh"

namespace hb {

// ------------------------------------------------------------------ forward block
template <int NT, int ACT>
__global__ void __launch_bounds__(256) linear_ln_fwd_kernel(const float* __restrict__ X, int ldx,
                                                            const float* __restrict__ WT, const float* __restrict__ bias,
                                                            const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                            float* __restrict__ Z, float* __restrict__ Y,
                                                            float* __restrict__ stats, int64_t M, int N, int Kred) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmSmem<NT>& s = *reinterpret_cast<GemmSmem<NT>*>(smem_raw);
  constexpr int NCH = NT / 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  float acc[4][NT / 16];
  gemm_mainloop<NT>(X, ldx, WT, N, M, Kred, N, row0, s, acc);

  float4 bv[NCH], gw[NCH], gb[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    int n = c * 64 + tx * 4;
    bv[c] = gw[c] = gb[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
      bv[c] = *reinterpret_cast<const float4*>(bias + n);
      gw[c] = *reinterpret_cast<const float4*>(lnw + n);
      gb[c] = *reinterpret_cast<const float4*>(lnb + n);
    }
  }
  const float inv_n = 1.f / (float)N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int n = c * 64 + tx * 4;
      float z0 = acc[i][c * 4 + 0] + bv[c].x, z1 = acc[i][c * 4 + 1] + bv[c].y;
      float z2 = acc[i][c * 4 + 2] + bv[c].z, z3 = acc[i][c * 4 + 3] + bv[c].w;
      if (Z != nullptr && row < M && n < N) *reinterpret_cast<float4*>(Z + row * N + n) = make_float4(z0, z1, z2, z3);
      bool ok = n < N;
      acc[i][c * 4 + 0] = ok ? act_fwd<ACT>(z0) : 0.f;
      acc[i][c * 4 + 1] = ok ? act_fwd<ACT>(z1) : 0.f;
      acc[i][c * 4 + 2] = ok ? act_fwd<ACT>(z2) : 0.f;
      acc[i][c * 4 + 3] = ok ? act_fwd<ACT>(z3) : 0.f;
      sum += acc[i][c * 4 + 0] + acc[i][c * 4 + 1] + acc[i][c * 4 + 2] + acc[i][c * 4 + 3];
    }
    const float mean = half_warp_sum(sum) * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c * 64 + tx * 4 < N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { float dlt = acc[i][c * 4 + j] - mean; sq = fmaf(dlt, dlt, sq); }
      }
    }
    const float var = half_warp_sum(sq) * inv_n;
    const float rstd = rsqrtf(var + 1e-5f);
    if (row < M) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        int n = c * 64 + tx * 4;
        if (n < N) {
          float4 y;
          y.x = (acc[i][c * 4 + 0] - mean) * rstd * gw[c].x + gb[c].x;
          y.y = (acc[i][c * 4 + 1] - mean) * rstd * gw[c].y + gb[c].y;
          y.z = (acc[i][c * 4 + 2] - mean) * rstd * gw[c].z + gb[c].z;
          y.w = (acc[i][c * 4 + 3] - mean) * rstd * gw[c].w + gb[c].w;
          *reinterpret_cast<float4*>(Y + row * N + n) = y;
        }
      }
      if (stats != nullptr && tx == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
    }
  }
}

// ------------------------------------------------------------------ backward block
// dYp = dZ [M,N] * W [N,Np];  g = dYp*gamma_p;  xh = (act(Zp)-mu)*rstd;
// dA = rstd*(g - mean(g) - xh*mean(g*xh));  dZp = dA*act'(Zp);  dgamma_p += sum_r dYp*xh; dbeta_p += sum_r dYp
template <int NT, int ACT>
__global__ void __launch_bounds__(256) dx_ln_bwd_kernel(const float* __restrict__ dZ, int N, const float* __restrict__ W,
                                                        const float* __restrict__ Zp, const float* __restrict__ stats_p,
                                                        const float* __restrict__ lnw_p, float* __restrict__ dZp,
                                                        float* __restrict__ g_lnw_p, float* __restrict__ g_lnb_p,
                                                        int64_t M, int Np) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GemmSmem<NT>& s = *reinterpret_cast<GemmSmem<NT>*>(smem_raw);
  constexpr int NCH = NT / 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  float acc[4][NT / 16];
  gemm_mainloop<NT>(dZ, N, W, Np, M, N, Np, row0, s, acc);

  float4 gw[NCH];
  float cg[NCH * 4], cb[NCH * 4];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    int n = c * 64 + tx * 4;
    gw[c] = n < Np ? *reinterpret_cast<const float4*>(lnw_p + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) cg[c * 4 + j] = cb[c * 4 + j] = 0.f;
  }
  const float inv_n = 1.f / (float)Np;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    const bool rok = row < M;
    float mu = 0.f, rstd = 0.f;
    if (rok) { mu = stats_p[row * 2]; rstd = stats_p[row * 2 + 1]; }
    float xh[NCH * 4], dact[NCH * 4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      int n = c * 64 + tx * 4;
      float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      bool ok = rok && n < Np;
      if (ok) z = *reinterpret_cast<const float4*>(Zp + row * Np + n);
      float zz[4] = {z.x, z.y, z.z, z.w};
      float gg[4] = {gw[c].x, gw[c].y, gw[c].z, gw[c].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float dy = ok ? acc[i][c * 4 + j] : 0.f;
        float x = ok ? (act_fwd<ACT>(zz[j]) - mu) * rstd : 0.f;
        xh[c * 4 + j] = x;
        dact[c * 4 + j] = act_bwd<ACT>(zz[j]);
        cg[c * 4 + j] = fmaf(dy, x, cg[c * 4 + j]);
        cb[c * 4 + j] += dy;
        float g = dy * gg[j];
        acc[i][c * 4 + j] = g;
        s1 += g;
        s2 = fmaf(g, x, s2);
      }
    }
    const float m1 = half_warp_sum(s1) * inv_n, m2 = half_warp_sum(s2) * inv_n;
    if (rok) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        int n = c * 64 + tx * 4;
        if (n < Np) {
          float4 o;
          o.x = rstd * (acc[i][c * 4 + 0] - m1 - xh[c * 4 + 0] * m2) * dact[c * 4 + 0];
          o.y = rstd * (acc[i][c * 4 + 1] - m1 - xh[c * 4 + 1] * m2) * dact[c * 4 + 1];
          o.z = rstd * (acc[i][c * 4 + 2] - m1 - xh[c * 4 + 2] * m2) * dact[c * 4 + 2];
          o.w = rstd * (acc[i][c * 4 + 3] - m1 - xh[c * 4 + 3] * m2) * dact[c * 4 + 3];
          *reinterpret_cast<float4*>(dZp + row * Np + n) = o;
        }
      }
    }
  }
  // column sums over the tile's 64 rows: lanes l and l^16 share columns -> shuffle, then 8 warps via smem
  float* red = reinterpret_cast<float*>(smem_raw);  // [8][2][NT]; mainloop finished with a barrier
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int q = 0; q < NCH * 4; ++q) {
    cg[q] += __shfl_xor_sync(0xffffffffu, cg[q], 16);
    cb[q] += __shfl_xor_sync(0xffffffffu, cb[q], 16);
  }
  if (lane < 16) {
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        red[(warp * 2 + 0) * NT + c * 64 + tx * 4 + j] = cg[c * 4 + j];
        red[(warp * 2 + 1) * NT + c * 64 + tx * 4 + j] = cb[c * 4 + j];
      }
  }
  __syncthreads();
  for (int n = tid; n < Np; n += 256) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[(w * 2 + 0) * NT + n]; b += red[(w * 2 + 1) * NT + n]; }
    atomicAdd(g_lnw_p + n, a);
    atomicAdd(g_lnb_p + n, b);
  }
}

// ------------------------------------------------------------------ weight gradient
// dW[n][k] += sum_{r in [m0,m1)} dZ[r][n] * X[r][k];  db[n] += sum_r dZ[r][n]
// CTA owns a 128(n) x KB(k) block over a slice of rows; threads TNX x TKX with float4 chunks.
template <int KB>
__global__ void __launch_bounds__(256) dw_accum_kernel(const float* __restrict__ dZ, int N, const float* __restrict__ X,
                                                       int ldx, int K, float* __restrict__ dW, float* __restrict__ db,
                                                       int64_t M, int64_t rows_per_cta) {
  constexpr int KCH = KB == 128 ? 2 : 1;  // float4 k-chunks per thread
  constexpr int NCH = KB == 128 ? 2 : 1;  // float4 n-chunks per thread
  constexpr int TKX = KB / (4 * KCH);     // threads along k: 16 / 8
  constexpr int TNX = 256 / TKX;          // threads along n: 16 / 32
  constexpr int RT = 32;                  // rows per smem tile
  __shared__ __align__(16) float zs[RT][128];
  __shared__ __align__(16) float xs[RT][KB];
  const int tid = threadIdx.x, tk = tid % TKX, tn = tid / TKX;
  const int n0 = blockIdx.y * 128, k0 = blockIdx.z * KB;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t m1 = m0 + rows_per_cta < M ? m0 + rows_per_cta : M;
  float acc[NCH * 4][KCH * 4];
  float bsum[NCH * 4];
#pragma unroll
  for (int i = 0; i < NCH * 4; ++i) {
    bsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < KCH * 4; ++j) acc[i][j] = 0.f;
  }
  for (int64_t r0 = m0; r0 < m1; r0 += RT) {
    // stage tiles (zero-filled outside the matrix)
#pragma unroll
    for (int q = 0; q < RT * 128 / 4 / 256; ++q) {
      int f = tid + q * 256;
      int r = f / 32, n4 = (f % 32) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < m1 && n0 + n4 < N) v = *reinterpret_cast<const float4*>(dZ + (r0 + r) * N + n0 + n4);
      *reinterpret_cast<float4*>(&zs[r][n4]) = v;
    }
#pragma unroll
    for (int q = 0; q < RT * KB / 4 / 256; ++q) {
      int f = tid + q * 256;
      int r = f / (KB / 4), k4 = (f % (KB / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < m1 && k0 + k4 < ldx) v = *reinterpret_cast<const float4*>(X + (r0 + r) * ldx + k0 + k4);
      *reinterpret_cast<float4*>(&xs[r][k4]) = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < RT; ++r) {
      float a[NCH * 4], b[KCH * 4];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float4 v = *reinterpret_cast<const float4*>(&zs[r][c * (TNX * 4) + tn * 4]);
        a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < KCH; ++c) {
        float4 v = *reinterpret_cast<const float4*>(&xs[r][c * (TKX * 4) + tk * 4]);
        b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < NCH * 4; ++i) {
        bsum[i] += a[i];
#pragma unroll
        for (int j = 0; j < KCH * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int ci = 0; ci < NCH; ++ci)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int n = n0 + ci * (TNX * 4) + tn * 4 + i;
      if (n >= N) continue;
#pragma unroll
      for (int cj = 0; cj < KCH; ++cj)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int k = k0 + cj * (TKX * 4) + tk * 4 + j;
          if (k < K) atomicAdd(dW + (int64_t)n * K + k, acc[ci * 4 + i][cj * 4 + j]);
        }
      if (db != nullptr && tk == 0 && blockIdx.z == 0) atomicAdd(db + n, bsum[ci * 4 + i]);
    }
}

// ------------------------------------------------------------------ launchers
template <int NT>
static int launch_fwd_nt(int act, const float* X, int ldx, const float* WT, const float* bias, const float* lnw,
                         const float* lnb, float* Z, float* Y, float* stats, int64_t M, int N, int Kred, cudaStream_t st) {
  const size_t smem = sizeof(GemmSmem<NT>);
  dim3 grid((unsigned)ceil_div64(M, BM));
#define HB_FWD_CASE(A)                                                                                          \
  case A: {                                                                                                     \
    auto kern = linear_ln_fwd_kernel<NT, A>;                                                                    \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    kern<<<grid, 256, smem, st>>>(X, ldx, WT, bias, lnw, lnb, Z, Y, stats, M, N, Kred);                         \
  } break;
  switch (act) {
    HB_FWD_CASE(HB_ACT_RELU) HB_FWD_CASE(HB_ACT_TANH) HB_FWD_CASE(HB_ACT_SIGMOID) HB_FWD_CASE(HB_ACT_LEAKY_RELU)
    HB_FWD_CASE(HB_ACT_SELU) HB_FWD_CASE(HB_ACT_HARDSWISH) HB_FWD_CASE(HB_ACT_IDENTITY)
    default: set_error("activation %d", act); return HB_ERR_UNSUPPORTED;
  }
#undef HB_FWD_CASE
  HB_LAUNCH_DONE(st, shape_label("linear_ln_fwd", M, N, Kred));
  return HB_OK;
}

int launch_linear_ln_fwd(int act, const float* X, int ldx, const float* WT, const float* bias, const float* lnw,
                         const float* lnb, float* Z, float* Y, float* stats, int64_t M, int N, int Kred, cudaStream_t st) {
  if (M <= 0) return HB_OK;
  if (N <= 64) return launch_fwd_nt<64>(act, X, ldx, WT, bias, lnw, lnb, Z, Y, stats, M, N, Kred, st);
  if (N <= 128) return launch_fwd_nt<128>(act, X, ldx, WT, bias, lnw, lnb, Z, Y, stats, M, N, Kred, st);
  return launch_fwd_nt<256>(act, X, ldx, WT, bias, lnw, lnb, Z, Y, stats, M, N, Kred, st);
}

template <int NT>
static int launch_dx_nt(int act, const float* dZ, int N, const float* W, const float* Zp, const float* stats_p,
                        const float* lnw_p, float* dZp, float* g_lnw_p, float* g_lnb_p, int64_t M, int Np, cudaStream_t st) {
  const size_t smem = sizeof(GemmSmem<NT>);
  dim3 grid((unsigned)ceil_div64(M, BM));
#define HB_DX_CASE(A)                                                                                           \
  case A: {                                                                                                     \
    auto kern = dx_ln_bwd_kernel<NT, A>;                                                                        \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    kern<<<grid, 256, smem, st>>>(dZ, N, W, Zp, stats_p, lnw_p, dZp, g_lnw_p, g_lnb_p, M, Np);                  \
  } break;
  switch (act) {
    HB_DX_CASE(HB_ACT_RELU) HB_DX_CASE(HB_ACT_TANH) HB_DX_CASE(HB_ACT_SIGMOID) HB_DX_CASE(HB_ACT_LEAKY_RELU)
    HB_DX_CASE(HB_ACT_SELU) HB_DX_CASE(HB_ACT_HARDSWISH) HB_DX_CASE(HB_ACT_IDENTITY)
    default: set_error("activation %d", act); return HB_ERR_UNSUPPORTED;
  }
#undef HB_DX_CASE
  HB_LAUNCH_DONE(st, shape_label("dx_ln_bwd", M, Np, N));
  return HB_OK;
}

int launch_dx_ln_bwd(int act, const float* dZ, int N, const float* W, const float* Zp, const float* stats_p,
                     const float* lnw_p, float* dZp, float* g_lnw_p, float* g_lnb_p, int64_t M, int Np, cudaStream_t st) {
  if (M <= 0) return HB_OK;
  if (Np <= 64) return launch_dx_nt<64>(act, dZ, N, W, Zp, stats_p, lnw_p, dZp, g_lnw_p, g_lnb_p, M, Np, st);
  if (Np <= 128) return launch_dx_nt<128>(act, dZ, N, W, Zp, stats_p, lnw_p, dZp, g_lnw_p, g_lnb_p, M, Np, st);
  return launch_dx_nt<256>(act, dZ, N, W, Zp, stats_p, lnw_p, dZp, g_lnw_p, g_lnb_p, M, Np, st);
}

int launch_dw_accum(const float* dZ, int N, const float* X, int ldx, int K, float* dW, float* db, int64_t M,
                    cudaStream_t st) {
  if (M <= 0) return HB_OK;
  const int nb = (N + 127) / 128;
  const bool small = K <= 32;
  const int kb = small ? 1 : (K + 127) / 128;
  // ~2 CTAs per SM overall; at least 64 rows per CTA so the atomics stay a small fraction
  int64_t splits = (2 * 148 + nb * kb - 1) / (nb * kb);
  int64_t rows_per = ceil_div64(M, splits);
  rows_per = (rows_per + 31) / 32 * 32;
  if (rows_per < 64) rows_per = 64;
  splits = ceil_div64(M, rows_per);
  dim3 grid((unsigned)splits, nb, kb);
  if (small) dw_accum_kernel<32><<<grid, 256, 0, st>>>(dZ, N, X, ldx, K, dW, db, M, rows_per);
  else dw_accum_kernel<128><<<grid, 256, 0, st>>>(dZ, N, X, ldx, K, dW, db, M, rows_per);
  HB_LAUNCH_DONE(st, shape_label("dw_accum", M, N, K));
  return HB_OK;
}

}  // namespace hb
