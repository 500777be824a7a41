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

// ---- rollout inference (fused_act_kernel): every actor and the critic of one rollout step in ONE launch
constexpr int ACT_MAX_NETS = 24;
struct ActNet {
  const float* prep; const float* obs; const float* avail;
  float* out0;                 // actions [rows, ad] | values [rows]
  float* out1;                 // log-probs [rows, ad] | unused
  unsigned long long seed;
  long long rows;
  int in_dim, out, head, K0p, nch0;
  int o_w0, o_w1, o_hw, o_b0, o_b1, o_bh, o_sc, o_ls;   // float offsets into prep (PrepLayout fz_* / log_std)
  float std_x, std_y;
  int tile0, stage;            // first CTA of this net; 1 = full tiles' inputs arrive by TMA bulk copy
};
struct ActArgs {
  int n_nets, H, act, deterministic, K0p_max, pad_;
  unsigned long long offset;
  const unsigned long long* offset_base;
  ActNet net[ACT_MAX_NETS];
};

}  // namespace fz

bool fused_enabled();
// all nets of one rollout step (hb_rollout_collect) / a single net (hb_policy_act, hb_value_forward) on the fused path;
// *handled = false: some net is outside the fused kernel's shapes -> caller falls back to the FP32 kernels
struct hb_collect_args;
int launch_fused_act(fz::ActArgs& A, cudaStream_t st);
int fused_act_fill(fz::ActNet* n, const hb_net_desc* d, const float* prepared, const float* obs, const float* avail, float* out0,
                   float* out1, unsigned long long seed, long long rows, bool* ok);
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
