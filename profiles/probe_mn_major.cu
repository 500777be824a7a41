// Stand-alone probe (not part of the library): do tcgen05 kind::tf32 MMAs accept the canonical no-swizzle shared-memory
// image of a row-major tile  IMG[row/8][chunk][row%8][4 floats]  as an MN-major operand, i.e. can the SAME image that
// serves  Y = X W^T  (K = features, K-major) also serve  G = X^T D  (K = rows) without a transposed re-staging?
// Tries both assignments of the descriptor's LBO / SBO fields for A and for B and prints the max error of each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe_mn profiles/probe_mn_major.cu && /tmp/probe_mn
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

constexpr int ROWS = 128, FEATS = 128, NJ = 16;

struct Smem {
  float x[ROWS * FEATS];   // [row/8][32 chunks][row%8][4]   64 KB
  float d[ROWS * NJ];      // [row/8][4 chunks][row%8][4]     8 KB
  uint64_t done;
  uint32_t tmem;
};

// variant bit 0: A fields swapped, bit 1: B fields swapped
__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ X, const float* __restrict__ D, float* __restrict__ G, int variant) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < ROWS * FEATS; i += 128) {
    const int r = i / FEATS, f = i % FEATS;
    s.x[((r >> 3) * 32 + (f >> 2)) * 32 + (r & 7) * 4 + (f & 3)] = X[i];
  }
  for (int i = tid; i < ROWS * NJ; i += 128) {
    const int r = i / NJ, j = i % NJ;
    s.d[((r >> 3) * 4 + (j >> 2)) * 32 + (r & 7) * 4 + (j & 3)] = D[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s.done)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s.tmem;
  // D fp32, A/B tf32, A and B MN-major (bits 15, 16), N = 16, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NJ >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (tid == 0) {
    for (int ks = 0; ks < ROWS / 8; ++ks) {  // one 8-row group per MMA k-step
      const uint32_t a0 = smem_u32(s.x) + ks * 32 * 128, b0 = smem_u32(s.d) + ks * 4 * 128;
      // A: m-groups (4 features) 128 B apart, k-groups (8 rows) 4096 B apart.  B: n-groups 128 B, k-groups 512 B.
      const uint64_t da = (variant & 1) ? umma_desc(a0, 128, 4096) : umma_desc(a0, 4096, 128);
      const uint64_t db = (variant & 2) ? umma_desc(b0, 128, 512) : umma_desc(b0, 512, 128);
      umma_tf32(tmem, da, db, idesc, ks != 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s.done)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&s.done)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < NJ; ++j) G[tid * NJ + j] = __uint_as_float(r[j]);   // lane = m = feature
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  static float X[ROWS * FEATS], D[ROWS * NJ], ref[FEATS * NJ], got[FEATS * NJ];
  srand(1);
  for (auto& v : X) v = (float)(rand() % 7 - 3);
  for (auto& v : D) v = (float)(rand() % 5 - 2);
  for (int f = 0; f < FEATS; ++f)
    for (int j = 0; j < NJ; ++j) {
      float a = 0.f;
      for (int r = 0; r < ROWS; ++r) a += X[r * FEATS + f] * D[r * NJ + j];
      ref[f * NJ + j] = a;
    }
  float *dX, *dD, *dG;
  cudaMalloc(&dX, sizeof X); cudaMalloc(&dD, sizeof D); cudaMalloc(&dG, sizeof got);
  cudaMemcpy(dX, X, sizeof X, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D, sizeof D, cudaMemcpyHostToDevice);
  const size_t smem = sizeof(Smem) + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int variant = 0; variant < 4; ++variant) {
    cudaMemset(dG, 0, sizeof got);
    probe<<<1, 128, smem>>>(dX, dD, dG, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(got, dG, sizeof got, cudaMemcpyDeviceToHost);
    float err = 0.f;
    for (int i = 0; i < FEATS * NJ; ++i) err = fmaxf(err, fabsf(got[i] - ref[i]));
    printf("variant %d (A %s, B %s): max |err| = %g   got[0..3] = %g %g %g %g   ref = %g %g %g %g\n", variant,
           (variant & 1) ? "LBO=m-group,SBO=k-group" : "LBO=k-group,SBO=m-group",
           (variant & 2) ? "LBO=n-group,SBO=k-group" : "LBO=k-group,SBO=n-group", err, got[0], got[1], got[2], got[3],
           ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
