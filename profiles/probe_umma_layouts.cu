// Stand-alone probe (not part of the library): which shared-memory operand layouts do tcgen05 MMAs accept for the
// fused update kernel?  One generic kernel; the host builds the operand images byte for byte, the descriptors and the
// expected result, and sweeps:
//   * kind::f16 (fp16 in, fp32 accumulate) and kind::tf32
//   * K-major and MN-major operands in the no-swizzle canonical layout of a ROW-WRITTEN tile
//       IMG[row/8][chunk][row%8][16 bytes]          (one thread = one row writes 16-byte chunks)
//     read (a) as a K-major operand (M/N index = row, K = feature) and (b) as an MN-major operand (M/N index =
//     feature, K = row) -- the same bytes must serve  Y = X W^T  and  dW = dZ^T X
//   * the same two views on a 128B-swizzled tile
//   * N = 16 / 144 (head GEMM, weight-gradient GEMM with a ones column), M = 128.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe_umma profiles/probe_umma_layouts.cu && /tmp/probe_umma
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Params {
  uint32_t a_bytes, b_bytes;     // image sizes
  uint32_t a_hi, b_hi;           // descriptor bits [32,64) (SBO, version, layout type)
  uint32_t a_lbo, b_lbo;         // descriptor bits [16,30) already shifted (<< 16)
  uint32_t a_step, b_step;       // start-address advance per k-step, bytes
  uint32_t idesc;
  int ksteps, n, tf32;
};

__global__ void __launch_bounds__(128, 1) probe(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B,
                                                float* __restrict__ D, Params p) {
  extern __shared__ __align__(1024) unsigned char raw[];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  uint8_t* sa = raw;
  uint8_t* sb = raw + ((p.a_bytes + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid * 16; i < p.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(A + i);
  for (uint32_t i = tid * 16; i < p.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(B + i);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    for (int ks = 0; ks < p.ksteps; ++ks) {
      const uint32_t a0 = smem_u32(sa) + ks * p.a_step, b0 = smem_u32(sb) + ks * p.b_step;
      const uint64_t da = ((uint64_t)p.a_hi << 32) | p.a_lbo | ((a0 & 0x3FFFFu) >> 4);
      const uint64_t db = ((uint64_t)p.b_hi << 32) | p.b_lbo | ((b0 & 0x3FFFFu) >> 4);
      const uint32_t acc = ks != 0;
      if (p.tf32)
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(p.idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(p.idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&done)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < p.n; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[tid * p.n + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// ---- host side --------------------------------------------------------------------------------------------------
enum Swz { NOSWZ = 0, SW128 = 2 };

// Byte offset of element (row r, feature f) of a ROW-WRITTEN tile with `feats` features of `esz` bytes.
//  NOSWZ: [r/8][f chunk (16 B)][r%8][16 B]
//  SW128: [f block (128 B worth)][r][128 B], 16-byte chunk index XOR (r % 8)
static size_t img_off(int swz, int r, int f, int rows, int feats, int esz) {
  const int per16 = 16 / esz;
  if (swz == NOSWZ) {
    const int nch = feats / per16;
    return ((size_t)(r / 8) * nch + f / per16) * 128 + (r % 8) * 16 + (size_t)(f % per16) * esz;
  }
  const int per128 = 128 / esz;
  const int blk = f / per128, fi = f % per128;
  const int ch = (fi / per16) ^ (r % 8);
  return (size_t)blk * rows * 128 + (size_t)r * 128 + ch * 16 + (size_t)(fi % per16) * esz;
}
static size_t img_bytes(int rows, int feats, int esz) { return (size_t)rows * feats * esz; }

struct Operand {
  std::vector<uint8_t> img;
  uint32_t hi, lbo, step;
};

// view = 0: K-major (MN index = row, K = feature); view = 1: MN-major (MN index = feature, K = row).
// swap = exchange the LBO / SBO fields (to find out which is which in the no-swizzle MN-major case).
static Operand make_operand(const std::vector<float>& X, int rows, int feats, bool tf32, int swz, int view, bool swap,
                            int kstep_elems) {
  const int esz = tf32 ? 4 : 2;
  Operand o;
  o.img.assign(img_bytes(rows, feats, esz), 0);
  for (int r = 0; r < rows; ++r)
    for (int f = 0; f < feats; ++f) {
      const size_t off = img_off(swz, r, f, rows, feats, esz);
      if (tf32) memcpy(&o.img[off], &X[(size_t)r * feats + f], 4);
      else { __half h = __float2half(X[(size_t)r * feats + f]); memcpy(&o.img[off], &h, 2); }
    }
  uint32_t lbo = 0, sbo = 0;
  const int per16 = 16 / esz, per128 = 128 / esz;
  if (swz == NOSWZ) {
    const uint32_t row_group = (uint32_t)(feats / per16) * 128;   // bytes between 8-row groups
    if (view == 0) { lbo = 128; sbo = row_group; o.step = (uint32_t)(kstep_elems / per16) * 128; }
    else { sbo = 128; lbo = row_group; o.step = (uint32_t)(kstep_elems / 8) * row_group; }
  } else {
    const uint32_t blk = (uint32_t)rows * 128;                      // bytes between 128-byte feature blocks
    if (view == 0) { lbo = 16; sbo = 1024; o.step = (uint32_t)kstep_elems * esz; (void)blk; }  // k-steps stay inside one 128 B block here
    else { lbo = blk; sbo = 1024; o.step = (uint32_t)(kstep_elems / 8) * 1024; }
    (void)per128;
  }
  if (swap) { uint32_t t = lbo; lbo = sbo; sbo = t; }
  o.lbo = ((lbo >> 4) & 0x3FFFu) << 16;
  o.hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)swz << 29);
  return o;
}

static int run_case(const char* name, bool tf32, int swz, int a_view, int b_view, bool a_swap, bool b_swap, int N, int K) {
  // logical problem: D[m][n] = sum_k A(m,k) B(n,k), M = 128.
  // K-major operand: tile rows = MN index, tile features = K.  MN-major operand: tile rows = K, features = MN index.
  const int M = 128;
  const int kstep = tf32 ? 8 : 16;
  std::vector<float> Al((size_t)M * K), Bl((size_t)N * K);
  srand(7);
  for (auto& v : Al) v = (float)(rand() % 7 - 3);
  for (auto& v : Bl) v = (float)(rand() % 5 - 2);
  std::vector<float> At, Bt;   // row-written tiles
  int a_rows, a_feats, b_rows, b_feats;
  if (a_view == 0) { a_rows = M; a_feats = K; At = Al; }
  else { a_rows = K; a_feats = M; At.resize((size_t)K * M); for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) At[(size_t)k * M + m] = Al[(size_t)m * K + k]; }
  if (b_view == 0) { b_rows = N; b_feats = K; Bt = Bl; }
  else { b_rows = K; b_feats = N; Bt.resize((size_t)K * N); for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bt[(size_t)k * N + n] = Bl[(size_t)n * K + k]; }
  if (swz == SW128) {
    const int per128 = tf32 ? 32 : 64;
    if (a_feats % per128 || b_feats % per128) { printf("%-58s skipped (feature count not a multiple of the 128 B block)\n", name); return 0; }
    if (a_view == 0 && a_feats > per128) { printf("%-58s skipped\n", name); return 0; }
    if (b_view == 0 && b_feats > per128) { printf("%-58s skipped\n", name); return 0; }
  }
  Operand oa = make_operand(At, a_rows, a_feats, tf32, swz, a_view, a_swap, kstep);
  Operand ob = make_operand(Bt, b_rows, b_feats, tf32, swz, b_view, b_swap, kstep);
  Params p;
  p.a_bytes = (uint32_t)oa.img.size(); p.b_bytes = (uint32_t)ob.img.size();
  p.a_hi = oa.hi; p.b_hi = ob.hi; p.a_lbo = oa.lbo; p.b_lbo = ob.lbo; p.a_step = oa.step; p.b_step = ob.step;
  const uint32_t fmt = tf32 ? 2u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_view << 15) | ((uint32_t)b_view << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  p.ksteps = K / kstep; p.n = N; p.tf32 = tf32;
  uint8_t *dA, *dB; float* dD;
  cudaMalloc(&dA, p.a_bytes); cudaMalloc(&dB, p.b_bytes); cudaMalloc(&dD, (size_t)M * N * 4);
  cudaMemcpy(dA, oa.img.data(), p.a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, ob.img.data(), p.b_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)M * N * 4);
  const size_t smem = ((p.a_bytes + 1023) / 1024) * 1024 + ((p.b_bytes + 1023) / 1024) * 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dD, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s CUDA error %s\n", name, cudaGetErrorString(e)); return 2; }
  std::vector<float> got((size_t)M * N);
  cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0.0, mag = 0.0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0.0;
      for (int k = 0; k < K; ++k) ref += (double)Al[(size_t)m * K + k] * Bl[(size_t)n * K + k];
      err = fmax(err, fabs(ref - got[(size_t)m * N + n]));
      mag = fmax(mag, fabs(ref));
    }
  printf("%-58s max|err| = %-10g (max|ref| = %g)  %s\n", name, err, mag, err == 0.0 ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return err == 0.0 ? 0 : 1;
}

// 3-pass fp16 split accuracy:  x = hi + lo (both fp16), D = A_hi B_hi + A_lo B_hi + A_hi B_lo  vs float64.
static void split_accuracy() {
  const int M = 128, N = 128, K = 128, kstep = 16;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(11);
  for (auto& v : A) v = ((float)rand() / RAND_MAX - 0.5f) * 6.f;   // LayerNorm-sized activations
  for (auto& v : B) v = ((float)rand() / RAND_MAX - 0.5f) * 0.5f;  // weights
  const float sb = 64.f;                                            // power-of-two weight scale
  std::vector<float> Ah(A.size()), Alo(A.size()), Bh(B.size()), Blo(B.size());
  for (size_t i = 0; i < A.size(); ++i) { Ah[i] = __half2float(__float2half(A[i])); Alo[i] = __half2float(__float2half(A[i] - Ah[i])); }
  for (size_t i = 0; i < B.size(); ++i) { float w = B[i] * sb; Bh[i] = __half2float(__float2half(w)); Blo[i] = __half2float(__float2half(w - Bh[i])); }
  // one launch with K tripled: [A_hi | A_lo | A_hi] x [B_hi | B_hi | B_lo]
  std::vector<float> A3((size_t)M * 3 * K), B3((size_t)N * 3 * K);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { A3[(size_t)m * 3 * K + k] = Ah[(size_t)m * K + k]; A3[(size_t)m * 3 * K + K + k] = Alo[(size_t)m * K + k]; A3[(size_t)m * 3 * K + 2 * K + k] = Ah[(size_t)m * K + k]; }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { B3[(size_t)n * 3 * K + k] = Bh[(size_t)n * K + k]; B3[(size_t)n * 3 * K + K + k] = Bh[(size_t)n * K + k]; B3[(size_t)n * 3 * K + 2 * K + k] = Blo[(size_t)n * K + k]; }
  Operand oa = make_operand(A3, M, 3 * K, false, NOSWZ, 0, false, kstep);
  Operand ob = make_operand(B3, N, 3 * K, false, NOSWZ, 0, false, kstep);
  Params p;
  p.a_bytes = (uint32_t)oa.img.size(); p.b_bytes = (uint32_t)ob.img.size();
  p.a_hi = oa.hi; p.b_hi = ob.hi; p.a_lbo = oa.lbo; p.b_lbo = ob.lbo; p.a_step = oa.step; p.b_step = ob.step;
  p.idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  p.ksteps = 3 * K / kstep; p.n = N; p.tf32 = 0;
  uint8_t *dA, *dB; float* dD;
  cudaMalloc(&dA, p.a_bytes); cudaMalloc(&dB, p.b_bytes); cudaMalloc(&dD, (size_t)M * N * 4);
  cudaMemcpy(dA, oa.img.data(), p.a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, ob.img.data(), p.b_bytes, cudaMemcpyHostToDevice);
  const size_t smem = ((p.a_bytes + 1023) / 1024) * 1024 + ((p.b_bytes + 1023) / 1024) * 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dD, p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("split accuracy: CUDA error %s\n", cudaGetErrorString(e)); return; }
  std::vector<float> got((size_t)M * N);
  cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost);
  double e3 = 0.0, e32 = 0.0, e1 = 0.0, mag = 0.0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0.0, one = 0.0; float f32 = 0.f;
      for (int k = 0; k < K; ++k) {
        ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
        f32 = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], f32);
        one += (double)Ah[(size_t)m * K + k] * Bh[(size_t)n * K + k] / sb;
      }
      e3 = fmax(e3, fabs(ref - got[(size_t)m * N + n] / sb));
      e32 = fmax(e32, fabs(ref - f32));
      e1 = fmax(e1, fabs(ref - one));
      mag = fmax(mag, fabs(ref));
    }
  printf("fp16 hi/lo 3-pass vs float64: max|err| = %g;  fp32 fmaf chain: %g;  single fp16 pass: %g   (max|ref| = %g)\n", e3, e32, e1, mag);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main(int argc, char** argv) {
  // one case per process (an illegal descriptor faults the context): argv[1] = case index, "split", or "count"
  int bad = 0;
  struct C { const char* name; bool tf32; int swz, av, bv; bool as, bs; int N, K; };
  const C cases[] = {
      {"f16  noswz A:K  B:K   N128 K32 (sanity)", false, NOSWZ, 0, 0, false, false, 128, 32},
      {"f16  noswz A:K  B:K   N16  K128 (head)", false, NOSWZ, 0, 0, false, false, 16, 128},
      {"f16  noswz A:K  B:K   N144 K128", false, NOSWZ, 0, 0, false, false, 144, 128},
      {"f16  noswz A:MN B:K   N128 K32", false, NOSWZ, 1, 0, false, false, 128, 32},
      {"f16  noswz A:K  B:MN  N128 K32", false, NOSWZ, 0, 1, false, false, 128, 32},
      {"f16  noswz A:MN B:MN  N144 K128 (dW)", false, NOSWZ, 1, 1, false, false, 144, 128},
      {"f16  noswz A:MN B:MN  N16 K128 (dWhead)", false, NOSWZ, 1, 1, false, false, 16, 128},
      {"f16  noswz A:K  B:MN  N128 K16 (dY from dlogit)", false, NOSWZ, 0, 1, false, false, 128, 16},
      {"tf32 noswz A:K  B:K   N128 K32 (sanity)", true, NOSWZ, 0, 0, false, false, 128, 32},
      {"tf32 noswz A:MN B:K   N128 K32", true, NOSWZ, 1, 0, false, false, 128, 32},
      {"tf32 noswz A:K  B:MN  N128 K32", true, NOSWZ, 0, 1, false, false, 128, 32},
      {"tf32 noswz A:MN B:MN  N128 K128", true, NOSWZ, 1, 1, false, false, 128, 128},
      {"f16  sw128 A:K  B:K   N128 K64", false, SW128, 0, 0, false, false, 128, 64},
      {"f16  sw128 A:MN B:K   N128 K64", false, SW128, 1, 0, false, false, 128, 64},
      {"f16  sw128 A:K  B:MN  N128 K64", false, SW128, 0, 1, false, false, 128, 64},
      {"f16  sw128 A:MN B:MN  N128 K128", false, SW128, 1, 1, false, false, 128, 128},
  };
  const int ncases = (int)(sizeof(cases) / sizeof(cases[0]));
  if (argc < 2 || !strcmp(argv[1], "count")) { printf("%d\n", ncases); return 0; }
  if (!strcmp(argv[1], "split")) { split_accuracy(); return 0; }
  const int i = atoi(argv[1]);
  if (i < 0 || i >= ncases) return 3;
  const C& c = cases[i];
  bad = run_case(c.name, c.tf32, c.swz, c.av, c.bv, c.as, c.bs, c.N, c.K);
  return bad;
}
