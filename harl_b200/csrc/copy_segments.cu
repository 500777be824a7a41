// One launch that copies up to HB_COPY_MAX_SEGS contiguous device segments (sm_100a).
//
// Used by the batched envs' step_into (harl_b200/envs/synthetic.py): per rollout step an env hands the runner A
// observation blocks, the shared state, the rewards and (Discrete) A availability masks, each of which lands in its
// own rollout-buffer slot (the reference copies them one by one in OnPolicyBaseRunner.insert,
// harl/runners/on_policy_base_runner.py:340-415).  Separate copy kernels are launch-bound (2-3 us each for < 1 MB)
// and torch's multi-tensor apply spends 18 us on these ~4 MB; here every CTA walks the concatenation of all segments
// in 16-byte vectors.  Either side of a segment may be pinned host memory (unified addressing): a host-resident env's step
// outputs then cross PCIe inside this one kernel instead of through three DMA set-ups and a device-side scatter, and the
// sampled actions go back the same way.
#include "common.cuh"

namespace hb {

struct CopySegs {
  void* dst[HB_COPY_MAX_SEGS];
  const void* src[HB_COPY_MAX_SEGS];
  long long end_v[HB_COPY_MAX_SEGS];   // exclusive prefix end of this segment, in 16-byte vectors
  long long tail_w[HB_COPY_MAX_SEGS];  // trailing 4-byte words (size % 16) copied by the first threads of CTA 0
  int n;
};

__global__ void __launch_bounds__(256) copy_segments_kernel(const __grid_constant__ CopySegs S) {
  const long long total = S.end_v[S.n - 1];
  const long long stride = (long long)gridDim.x * 256;
  int seg = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) {
    while (i >= S.end_v[seg]) ++seg;
    const long long off = i - (seg ? S.end_v[seg - 1] : 0);
    reinterpret_cast<uint4*>(S.dst[seg])[off] = __ldcv(reinterpret_cast<const uint4*>(S.src[seg]) + off);   // .cv: a source may be host memory the CPU rewrites every step
  }
  if (blockIdx.x == 0) {
    for (int s = 0; s < S.n; ++s) {
      const long long nv = S.end_v[s] - (s ? S.end_v[s - 1] : 0);
      if (threadIdx.x < S.tail_w[s])
        reinterpret_cast<uint32_t*>(S.dst[s])[nv * 4 + threadIdx.x] = __ldcv(reinterpret_cast<const uint32_t*>(S.src[s]) + nv * 4 + threadIdx.x);
    }
  }
}

}  // namespace hb

extern "C" int hb_copy_segments(const hb_copy_seg* segs, int32_t n, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(segs && n > 0 && n <= HB_COPY_MAX_SEGS, "1..HB_COPY_MAX_SEGS segments");
  CopySegs S;
  S.n = n;
  long long acc = 0;
  for (int i = 0; i < n; ++i) {
    HB_CHECK_ARG(segs[i].dst && segs[i].src && segs[i].bytes >= 0, "NULL segment");
    HB_CHECK_ARG(((uintptr_t)segs[i].dst & 15) == 0 && ((uintptr_t)segs[i].src & 15) == 0 && segs[i].bytes % 4 == 0,
                 "segments must start on 16-byte boundaries and hold whole 4-byte words");
    S.dst[i] = segs[i].dst;
    S.src[i] = segs[i].src;
    acc += segs[i].bytes / 16;
    S.end_v[i] = acc;
    S.tail_w[i] = (segs[i].bytes % 16) / 4;
  }
  if (acc == 0) {
    bool any = false;
    for (int i = 0; i < n; ++i) any |= S.tail_w[i] != 0;
    if (!any) return HB_OK;
  }
  long long ctas = (acc + 256 * 4 - 1) / (256 * 4);   // >= 4 vectors per thread before adding CTAs
  ctas = ctas < 1 ? 1 : (ctas > 148 * 4 ? 148 * 4 : ctas);
  copy_segments_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(S);
  HB_LAUNCH_DONE((cudaStream_t)stream, "hb_copy_segments");
  return HB_OK;
}
