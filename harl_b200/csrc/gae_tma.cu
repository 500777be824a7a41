// GAE / return scan with TMA tile staging (sm_100a): the GAE branch of hb_gae_returns for the common case
// (use_gae, 16-byte aligned buffers, C % 4 == 0).  Same arithmetic, bit for bit, as gae_tiled_kernel in gae.cu
// (reference: harl/common/buffers/on_policy_critic_buffer_ep.py:97-200, harl/runners/on_policy_ha_runner.py:26-33).
//
// Experiment (opt-in, see launch_gae_tma): instead of one 16-byte cp.async per thread-instruction, one
// elected thread issues four 2-D TMA box loads per time chunk (cp.async.bulk.tensor.2d, box = CW columns x BR rows of
// rewards / value_preds / masks / bad_masks) that land dense [row][CW] tiles in shared memory and signal the chunk's
// mbarrier; no per-element load instructions at all.  Chunks are issued last-first: the backward recurrence starts
// on the last chunk while earlier ones are in flight.
#include <cuda.h>
#include <stdlib.h>

#include <map>
#include <tuple>

#include "common.cuh"

namespace hb {

namespace {

struct VNc { float mean, std; int on; };
__device__ __forceinline__ VNc vn_ld(const float* __restrict__ vn) {
  VNc c;
  c.on = vn != nullptr;
  c.mean = 0.f;
  c.std = 1.f;
  if (c.on) {  // valuenorm.py:38-45,78-92
    float d = fmaxf(vn[2], 1e-5f);
    float m = __fdiv_rn(vn[0], d), msq = __fdiv_rn(vn[1], d);
    float var = fmaxf(__fsub_rn(msq, __fmul_rn(m, m)), 1e-2f);
    c.mean = m;
    c.std = __fsqrt_rn(var);
  }
  return c;
}
__device__ __forceinline__ float dn(const VNc& c, float v) { return c.on ? __fadd_rn(__fmul_rn(v, c.std), c.mean) : v; }

__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(saddr(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_box_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(saddr(dst)), "l"(map), "r"(x), "r"(y), "r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_row(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(saddr(dst)), "l"(src), "r"(bytes), "r"(saddr(bar)) : "memory");
}

constexpr int GT_CH = 4;   // time chunks

struct GaeMaps { CUtensorMap rew, val, msk, bad; };

// smem: four arrays of RA = GT_CH * BR + 2 rows x CW floats.  Rows [0, GT_CH*BR) are written by the boxes (rows >= T
// zero-filled: the maps declare T rows); the bootstrap row (next_value, masks[T], bad_masks[T]) lives in row RT = GT_CH*BR.
template <int CW>
__global__ void __launch_bounds__(256) gae_tma_kernel(const __grid_constant__ GaeMaps maps, float* __restrict__ value_preds,
                                                      const float* __restrict__ masks, const float* __restrict__ bad_masks,
                                                      const float* __restrict__ next_value, float* __restrict__ returns,
                                                      float* __restrict__ adv, int T, int BR, int64_t C, float gamma, float gl,
                                                      int ptl, const float* __restrict__ vn) {
  extern __shared__ __align__(128) float gsm[];
  __shared__ __align__(8) unsigned long long bars[GT_CH];
  const int RT = GT_CH * BR, RA = RT + 2;   // + bootstrap row + one pad row: every array starts 128-byte aligned
  float* s_v = gsm;
  float* s_m = s_v + (size_t)RA * CW;
  float* s_b = s_m + (size_t)RA * CW;
  float* s_w = s_b + (size_t)RA * CW;
  const int64_t c0 = (int64_t)blockIdx.x * CW;
  const int tid = threadIdx.x;
  const VNc vc = vn_ld(vn);
  const int cwv = (int)(C - c0 < CW ? C - c0 : CW);
  if (tid == 0) {
    for (int k = 0; k < GT_CH; ++k) bar_init(&bars[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned box_bytes = (unsigned)(CW * BR) * 4u;
    for (int k = GT_CH - 1; k >= 0; --k) {
      const unsigned extra = k == GT_CH - 1 ? 3u * (unsigned)cwv * 4u : 0u;
      bar_expect(&bars[k], 4u * box_bytes + extra);
      const int t0 = k * BR;
      tma_box_2d(s_v + (size_t)t0 * CW, &maps.val, (int)c0, t0, &bars[k]);
      tma_box_2d(s_w + (size_t)t0 * CW, &maps.rew, (int)c0, t0, &bars[k]);
      tma_box_2d(s_m + (size_t)t0 * CW, &maps.msk, (int)c0, t0, &bars[k]);
      tma_box_2d(s_b + (size_t)t0 * CW, &maps.bad, (int)c0, t0, &bars[k]);
      if (k == GT_CH - 1) {
        bulk_row(s_v + (size_t)RT * CW, next_value + c0, (unsigned)cwv * 4u, &bars[k]);
        bulk_row(s_m + (size_t)RT * CW, masks + (int64_t)T * C + c0, (unsigned)cwv * 4u, &bars[k]);
        bulk_row(s_b + (size_t)RT * CW, bad_masks + (int64_t)T * C + c0, (unsigned)cwv * 4u, &bars[k]);
      }
    }
  }
  __syncthreads();
  float g = 0.f;   // running GAE of this thread's column (threads < CW)
#pragma unroll 1
  for (int k = GT_CH - 1; k >= 0; --k) {
    bar_wait(&bars[k], 0);
    const int t0 = k * BR < T ? k * BR : T;
    const int t1 = (k + 1) * BR < T ? (k + 1) * BR : T;
    if (k == GT_CH - 1) {  // value_preds[-1] = next_value; denormalise the bootstrap row
      for (int f = tid; f < CW; f += 256) {
        const float nv = s_v[RT * CW + f];
        if (c0 + f < C) value_preds[(int64_t)T * C + c0 + f] = nv;
        s_v[RT * CW + f] = dn(vc, nv);
      }
    }
    for (int f = t0 * CW + tid; f < t1 * CW; f += 256) s_v[f] = dn(vc, s_v[f]);
    __syncthreads();
    for (int f = t0 * CW + tid; f < t1 * CW; f += 256) {
      const int t = f / CW, c = f % CW;
      const int n = (t + 1 == T ? RT : t + 1) * CW + c;
      // delta = r + gamma * V^[t+1] * m[t+1] - V^[t]
      s_w[f] = __fsub_rn(__fadd_rn(s_w[f], __fmul_rn(__fmul_rn(gamma, s_v[n]), s_m[n])), s_v[f]);
    }
    __syncthreads();
    if (tid < CW && c0 + tid < C) {
      for (int tb = t1 - 1; tb >= t0; tb -= 8) {
        float w[8], am[8], bb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int t = tb - i;
          w[i] = am[i] = 0.f;
          bb[i] = 1.f;
          if (t >= t0) {
            const int n = (t + 1 == T ? RT : t + 1) * CW + tid;
            w[i] = s_w[t * CW + tid];
            am[i] = __fmul_rn(gl, s_m[n]);
            if (ptl) bb[i] = s_b[n];
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int t = tb - i;
          if (t >= t0) {
            g = __fadd_rn(w[i], __fmul_rn(am[i], g));
            if (ptl) g = __fmul_rn(bb[i], g);
            s_w[t * CW + tid] = g;
          }
        }
      }
    }
    __syncthreads();
    for (int f = t0 * CW + tid; f < t1 * CW; f += 256) {
      const int t = f / CW, c = f % CW;
      if (c0 + c < C) {
        const float r = __fadd_rn(s_w[f], s_v[f]);
        returns[(int64_t)t * C + c0 + c] = r;
        if (adv) adv[(int64_t)t * C + c0 + c] = __fsub_rn(r, s_v[f]);
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

bool make_map(CUtensorMap* m, const float* base, int64_t C, int rows, int cw, int br) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)C * 4u};
  const cuuint32_t box[2] = {(cuuint32_t)cw, (cuuint32_t)br};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Returns true when it launched (the caller then skips the cp.async kernel).  *rc receives the status.
bool launch_gae_tma(const float* rewards, float* value_preds, const float* masks, const float* bad_masks,
                    const float* next_value, float* returns, float* advantages, int T, int64_t C, float gamma, float gl,
                    int ptl, const float* vn, cudaStream_t st, int* rc) {
  // Opt-in (HB_GAE_TMA=1).  Measured on B200 (profiles/gae_variants_r01.txt): 17.9 us per launch at [200, 4096] and
  // 129 us at [200, 65536], against 13.2 / 104 us for the 16-byte cp.async kernel in gae.cu -- the staging is not what
  // bounds this kernel (the per-chunk barriers around the serial recurrence are), so the cp.async kernel stays default.
  static const bool on = getenv("HB_GAE_TMA") && atoi(getenv("HB_GAE_TMA")) == 1;
  if (!on || T < GT_CH || C % 4 != 0 || C >= (1ll << 31)) return false;
  const int BR = ((T + GT_CH - 1) / GT_CH + 1) & ~1;   // even: every chunk of a 16-column tile starts 128-byte aligned
  if (BR > 256) return false;
  // tile width: 32 columns (128-byte rows) when that still fills the GPU, else 16
  int cw = ceil_div64(C, 32) >= 148 ? 32 : 16;
  static const int forced = getenv("HB_GAE_CW") ? atoi(getenv("HB_GAE_CW")) : 0;
  if (forced == 16 || forced == 32) cw = forced;
  const size_t smem = (size_t)4 * (GT_CH * BR + 2) * cw * sizeof(float);
  if (smem > 200 * 1024) return false;
  // tensor maps depend only on (pointers, T, C, tile): cache them (the buffers are allocated once per run)
  typedef std::tuple<const void*, const void*, const void*, const void*, int, int64_t, int> Key;
  static thread_local std::map<Key, GaeMaps> cache;
  const Key key(rewards, value_preds, masks, bad_masks, T, C, cw);
  auto it = cache.find(key);
  if (it == cache.end()) {
    GaeMaps m;
    if (!make_map(&m.rew, rewards, C, T, cw, BR) || !make_map(&m.val, value_preds, C, T, cw, BR) ||
        !make_map(&m.msk, masks, C, T, cw, BR) || !make_map(&m.bad, bad_masks, C, T, cw, BR))
      return false;
    if (cache.size() > 64) cache.clear();
    it = cache.emplace(key, m).first;
  }
  const unsigned grid = (unsigned)ceil_div64(C, cw);
  if (cw == 32) {
    cudaFuncSetAttribute(gae_tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gae_tma_kernel<32><<<grid, 256, smem, st>>>(it->second, value_preds, masks, bad_masks, next_value, returns, advantages, T, BR,
                                                C, gamma, gl, ptl, vn);
  } else {
    cudaFuncSetAttribute(gae_tma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gae_tma_kernel<16><<<grid, 256, smem, st>>>(it->second, value_preds, masks, bad_masks, next_value, returns, advantages, T, BR,
                                                C, gamma, gl, ptl, vn);
  }
  cudaError_t e = cudaGetLastError();
  *rc = e == cudaSuccess ? HB_OK : cuda_fail(e, "hb_gae_returns(tma)");
  if (e == cudaSuccess) note_launch("hb_gae_returns", st);
  return true;
}

}  // namespace hb
