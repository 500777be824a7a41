// Launcher interface of the trust-region (HATRPO) kernels (trpo.cu), used by the sequencing code in capi.cu.
#pragma once
#include "common.cuh"

namespace hb {

enum { TR_OLD = 0, TR_FVP = 1, TR_LS = 2 };

struct TrpoHeadArgs {
  const float* feat;      // [rows, h] trunk output of the parameters under evaluation
  int h, out;
  const float* hw;        // [out][h]
  const float* hbias;     // [out]
  const float* log_std;   // [out] (Box, current parameters)
  float std_x, std_y;
  int64_t rows;
  const int32_t* index;   // batch row -> buffer row (nullable)
  const float* avail;     // buffer-row indexed (Discrete, nullable)
  // OLD
  float* old_dist_out;    // [rows, out] normalised logits (Discrete) / means (Box), batch-row indexed
  // FVP
  const float* featd;     // [rows, h] tangent of the trunk output
  const float* hwd;       // [out][h] tangent of the head weights
  const float* hbd;       // [out]
  const float* old_dist;  // [rows, out]
  float inv_rows;         // 1 / (global number of rows): the KL is a plain mean over rows
  float* dfeat;           // [rows, h] receives dZ of the last trunk block
  float* g_hw; float* g_hbias;
  const float* ln_z; const float* ln_stats; const float* ln_w; float* g_ln_w; float* g_ln_b; int ln_act;
  int64_t part_delta, part_stride;
  // LS (buffer-row indexed)
  const float* actions; const float* old_logp; const float* adv; const float* factor; const float* active;
  const float* old_log_std;  // [out] (Box, parameters before the step)
  int use_active, agg_prod;
  double* scalars;        // += (sum ratio*factor*adv*w, sum entropy*w, sum ratio, sum KL(old || new))
};

int launch_tangent_prepare(const hb_net_desc* d, const ParamLayout& P, const PrepLayout& Q, const float* params,
                           const float* v, float* tprep, cudaStream_t st);
int launch_jvp_linear_ln(int act, const float* X, int ldx, const float* Xd, const float* WT, const float* WdT,
                         const float* bd, const float* lnw, const float* lnwd, const float* lnbd, const float* Z,
                         const float* stats, float* Yd, int64_t M, int N, int Kred, cudaStream_t st);
int launch_trpo_head(int head, int mode, const TrpoHeadArgs& a, cudaStream_t st);
int launch_cg_init(const float* b, float* x, float* r, float* p, float* state, int n, cudaStream_t st);
int launch_cg_step(float* p, const float* avp, float* x, float* r, float* state, int n, float tol, cudaStream_t st);
int launch_full_step(const float* x, const float* fx, const float* g, float kl_threshold, float* full, double* out3, int n,
                     cudaStream_t st);
int launch_apply_step(float* params, const float* params0, const float* full, float fraction, int n, cudaStream_t st);
int launch_vec_scale(float* x, float s, int n, cudaStream_t st);
int launch_fvp_finish(float* out, const float* v, const float* params, float damping, int n, int ls_off, int ls_n,
                      float std_x, cudaStream_t st);

}  // namespace hb
