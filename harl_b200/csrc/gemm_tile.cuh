// 64 x NT FP32 SIMT GEMM tile shared by gemm_simt.cu (forward / backward blocks) and trpo.cu (tangent blocks).
#pragma once
#include "common.cuh"

namespace hb {

constexpr int BM = 64;
constexpr int KC = 16;
constexpr int LDA_S = KC + 4;

template <int NT>
struct GemmSmem {
  float a[2][BM][LDA_S];
  float b[2][KC][NT];
};

// acc[i][c*4+j] += sum_k A[row0+ty*4+i][k] * B[k][c*64+tx*4+j]
// ZERO = false keeps the incoming accumulators (a second product summed into the same tile, trpo.cu)
template <int NT, bool ZERO = true>
__device__ __forceinline__ void gemm_mainloop(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                              int ldb, int64_t M, int Kred, int Nout, int64_t row0,
                                              GemmSmem<NT>& s, float (&acc)[4][NT / 16]) {
  constexpr int NCH = NT / 64;
  constexpr int BLD = KC * NT / 4 / 256;  // float4 per thread for the B tile
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (ZERO) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NT / 16; ++j) acc[i][j] = 0.f;
  }

  const int a_r = tid >> 2, a_k = (tid & 3) * 4;
  const int64_t a_row = row0 + a_r;
  float4 ra;
  float4 rb[BLD];
  auto load_tiles = [&](int k0) {
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_row < M && k0 + a_k < Kred) ra = *reinterpret_cast<const float4*>(A + a_row * lda + k0 + a_k);
#pragma unroll
    for (int q = 0; q < BLD; ++q) {
      int f = tid + q * 256;
      int kk = f / (NT / 4), n4 = (f % (NT / 4)) * 4;
      rb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < Kred && n4 < Nout) rb[q] = *reinterpret_cast<const float4*>(B + (int64_t)(k0 + kk) * ldb + n4);
    }
  };
  auto store_tiles = [&](int buf) {
    *reinterpret_cast<float4*>(&s.a[buf][a_r][a_k]) = ra;
#pragma unroll
    for (int q = 0; q < BLD; ++q) {
      int f = tid + q * 256;
      int kk = f / (NT / 4), n4 = (f % (NT / 4)) * 4;
      *reinterpret_cast<float4*>(&s.b[buf][kk][n4]) = rb[q];
    }
  };
  const int nk = (Kred + KC - 1) / KC;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * KC);
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(&s.a[buf][ty * 4 + i][k4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 bv[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) bv[c] = *reinterpret_cast<const float4*>(&s.b[buf][k4 + kk][c * 64 + tx * 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            acc[i][c * 4 + 0] = fmaf(a, bv[c].x, acc[i][c * 4 + 0]);
            acc[i][c * 4 + 1] = fmaf(a, bv[c].y, acc[i][c * 4 + 1]);
            acc[i][c * 4 + 2] = fmaf(a, bv[c].z, acc[i][c * 4 + 2]);
            acc[i][c * 4 + 3] = fmaf(a, bv[c].w, acc[i][c * 4 + 3]);
          }
        }
      }
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }
}

}  // namespace hb
