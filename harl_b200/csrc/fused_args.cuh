// Argument block and launchers of the fused update kernel (fused_update.cu), shared with capi.cu.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace hb {
namespace fz {

struct Args {
  // network
  int H, K0p, in_dim, out, act, nch0;
  const __half* img0; const __half* img1; const __half* img1b; const __half* imgh;
  const float* bias0; const float* bias1; const float* biash;
  const float* scales;                      // [3] power-of-two operand scales of W'_0, W'_1, W'_head
  const float* log_std; float std_x, std_y;
  // batch (buffer-row indexed through `index`)
  const float* obs; const int32_t* index; long long rows;
  int stage_obs;                            // full tiles' observation blocks arrive by TMA bulk copy (identity index, 16-byte aligned)
  const float* actions; const float* avail; const float* old_logp; const float* adv; const float* factor; const float* active;
  const float* value_preds; const float* returns; const float* vn_state;
  // hyper-parameters
  float clip, entropy_coef, huber_delta, vcoef;
  int use_active, use_clip, agg_prod, use_huber, use_clipped;
  // gradient mode outputs
  float* part; long long part_stride;
  int pw0, pb0, pw1, pb1, phw, phb, plogstd;
  double* scalars;
  // evaluate mode outputs
  float* logp_out; const float* logp_ref; float* factor_inout;
};

}  // namespace fz

bool fused_enabled();
void set_fused_enabled(int on);
inline int fused_max_slots() { return 148; }   // persistent grid: one CTA (one split-buffer slot) per SM
// mode 0 = forward + loss + backward into the split buffer, 1 = evaluate (log-probs / values, factor update)
int launch_fused_update(const hb_net_desc* d, const PrepLayout& Q, const ParamLayout& P, const float* prepared, fz::Args a, int mode,
                        int* grid_out, cudaStream_t st);
// slot sums -> grad with the loss normaliser, then the LayerNorm-affine unfolding
int launch_fused_finish(const hb_net_desc* d, const ParamLayout& P, const float* params, float* grad, const float* part, int slots,
                        long long stride, const double* norm3, double host_scale, cudaStream_t st);

int fused_timing_enable(int on);
int fused_timing_read(unsigned long long* out);

}  // namespace hb
