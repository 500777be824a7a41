// C-ABI entry points that sequence the kernels for one network over row chunks.
//
// Rows are processed in chunks of up to 2^20 (the workspace is sized for one chunk: ~3 KB per row at hidden 128,
// i.e. ~3 GB -- nothing against 180 GB of HBM3e).  Measured on B200 (profiles/chunk_sweep_r01.txt): L2-sized
// chunks of 32768 rows leave the GPU with 256 CTAs per launch -- 1.7 per SM, latency-bound at every kernel --
// and cost 94 ms per C2 update phase; whole-batch launches (819200 rows, 6400 CTAs) take 55 ms although the
// activations then stream through HBM.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace hb {

// Rows per kernel launch.  HB_CHUNK_ROWS overrides it for tuning runs (read once).
static int64_t chunk_rows_init() {
  const char* e = getenv("HB_CHUNK_ROWS");
  long long v = e ? atoll(e) : 0;
  return v >= 1024 ? (int64_t)v : (int64_t)1 << 20;
}
static const int64_t CHUNK_ROWS = chunk_rows_init();

struct Work {
  float* x0;
  float* Z[HB_MAX_LAYERS];
  float* Y[HB_MAX_LAYERS];
  float* stats[HB_MAX_LAYERS];
  float* dA;
  float* dB;
  float* dwpart;   // [tc_dw_splits()][param total] split buffer of the tensor-core dW kernel (gradient mode)
  int ptotal;
};

static int hmax_of(const PrepLayout& Q) {
  int m = 0;
  for (int l = 0; l < Q.n_layers; ++l) m = Q.n[l] > m ? Q.n[l] : m;
  return m;
}

static size_t work_floats(const PrepLayout& Q, int64_t ch, int mode, int ptotal) {
  size_t f = (size_t)ch * Q.kpad[0];
  const size_t hm = hmax_of(Q);
  if (mode == 0) return f + 2 * (size_t)ch * hm;
  for (int l = 0; l < Q.n_layers; ++l) f += 2 * (size_t)ch * Q.n[l] + (size_t)round_up((int)(2 * ch), 4);
  return f + 2 * (size_t)ch * hm + (size_t)tc_dw_splits() * ptotal;
}

static int carve(const PrepLayout& Q, int64_t ch, int mode, void* ws, size_t ws_bytes, Work* w, int ptotal = 0) {
  if (work_floats(Q, ch, mode, ptotal) * sizeof(float) > ws_bytes || ws == nullptr) {
    set_error("workspace too small: need %zu bytes, have %zu", work_floats(Q, ch, mode, ptotal) * sizeof(float), ws_bytes);
    return HB_ERR_WORKSPACE;
  }
  w->ptotal = ptotal;
  w->dwpart = nullptr;
  float* p = (float*)ws;
  const size_t hm = hmax_of(Q);
  w->x0 = p; p += (size_t)ch * Q.kpad[0];
  if (mode == 0) {
    w->dA = p; p += (size_t)ch * hm;
    w->dB = p;
    for (int l = 0; l < Q.n_layers; ++l) { w->Z[l] = nullptr; w->stats[l] = nullptr; w->Y[l] = (l & 1) ? w->dB : w->dA; }
    return HB_OK;
  }
  for (int l = 0; l < Q.n_layers; ++l) {
    w->Z[l] = p; p += (size_t)ch * Q.n[l];
    w->Y[l] = p; p += (size_t)ch * Q.n[l];
    w->stats[l] = p; p += (size_t)round_up((int)(2 * ch), 4);
  }
  w->dA = p; p += (size_t)ch * hm;
  w->dB = p; p += (size_t)ch * hm;
  w->dwpart = p;
  return HB_OK;
}

// feature norm + trunk for `rows` rows starting at buffer row c0 (or index + c0)
static int trunk_forward(const hb_net_desc* d, const PrepLayout& Q, const float* prep, const float* obs,
                         const int32_t* index, int64_t c0, int64_t rows, const Work& w, cudaStream_t st) {
  const float* o = index ? obs : obs + c0 * d->in_dim;
  const int32_t* idx = index ? index + c0 : nullptr;
  int rc = launch_feat_norm(o, d->in_dim, idx, rows, d->feature_norm, w.x0, Q.kpad[0], st);
  if (rc) return rc;
  const float* x = w.x0;
  int ldx = Q.kpad[0];
  for (int l = 0; l < Q.n_layers; ++l) {
    const int impl = gemm_impl();
    if (impl != 0)
      rc = launch_tc_linear_ln_fwd(impl == 1 ? 3 : 1, d->activation, x, ldx, prep + Q.tk[l], Q.tk_chunks[l],
                                   prep + Q.bias[l], prep + Q.lnw[l], prep + Q.lnb[l], w.Z[l], w.Y[l], w.stats[l], rows,
                                   Q.n[l], Q.kpad[l], st);
    else
      rc = launch_linear_ln_fwd(d->activation, x, ldx, prep + Q.wt[l], prep + Q.bias[l], prep + Q.lnw[l], prep + Q.lnb[l],
                                w.Z[l], w.Y[l], w.stats[l], rows, Q.n[l], Q.kpad[l], st);
    if (rc) return rc;
    x = w.Y[l];
    ldx = Q.n[l];
  }
  return HB_OK;
}

// backward through the trunk given d(loss)/d(features) in w.dA
static int trunk_backward(const hb_net_desc* d, const ParamLayout& P, const PrepLayout& Q, const float* params,
                          const float* prep, float* grad, int64_t rows, const Work& w, cudaStream_t st) {
  const int L = Q.n_layers;
  float* dcur = w.dA;
  float* dnext = w.dB;
  int rc = HB_OK;  // the LN + activation backward of the last block is fused into the head kernel (w.dA already holds dZ_L)
  const int impl = gemm_impl();
  const int passes = impl == 1 ? 3 : 1;
  for (int l = L - 1; l >= 1; --l) {
    if (impl != 0)
      rc = launch_tc_dw_accum(passes, dcur, Q.n[l], w.Y[l - 1], Q.n[l - 1], Q.k[l], w.dwpart + P.w[l], grad + P.b[l], rows,
                              w.ptotal, st);
    else
      rc = launch_dw_accum(dcur, Q.n[l], w.Y[l - 1], Q.n[l - 1], Q.k[l], grad + P.w[l], grad + P.b[l], rows, st);
    if (rc) return rc;
    if (impl != 0)
      rc = launch_tc_dx_ln_bwd(passes, d->activation, dcur, Q.n[l], prep + Q.tkt[l], Q.tkt_chunks[l], w.Z[l - 1],
                               w.stats[l - 1], prep + Q.lnw[l - 1], dnext, grad + P.lnw[l - 1], grad + P.lnb[l - 1], rows,
                               Q.n[l - 1], w.dwpart - grad, w.ptotal, st);
    else
      rc = launch_dx_ln_bwd(d->activation, dcur, Q.n[l], params + P.w[l], w.Z[l - 1], w.stats[l - 1], prep + Q.lnw[l - 1],
                            dnext, grad + P.lnw[l - 1], grad + P.lnb[l - 1], rows, Q.n[l - 1], st);
    if (rc) return rc;
    float* t = dcur; dcur = dnext; dnext = t;
  }
  if (impl != 0)
    return launch_tc_dw_accum(passes, dcur, Q.n[0], w.x0, Q.kpad[0], Q.k[0], w.dwpart + P.w[0], grad + P.b[0], rows, w.ptotal, st);
  return launch_dw_accum(dcur, Q.n[0], w.x0, Q.kpad[0], Q.k[0], grad + P.w[0], grad + P.b[0], rows, st);
}

static int check_net(const hb_net_desc* d, ParamLayout* P, PrepLayout* Q, hb_net_layout* L, int want_policy) {
  int rc = make_layouts(d, P, Q, L);
  if (rc) return rc;
  if (d->rnn_layers) { set_error("recurrent (GRU) networks are not implemented in this build"); return HB_ERR_UNSUPPORTED; }
  if (want_policy == 1 && d->head == HB_HEAD_VALUE) { set_error("expected a policy head"); return HB_ERR_INVALID; }
  if (want_policy == 0 && d->head != HB_HEAD_VALUE) { set_error("expected a value head"); return HB_ERR_INVALID; }
  return HB_OK;
}

static void head_base(const hb_net_desc* d, const PrepLayout& Q, const float* prep, HeadArgs* a) {
  memset(a, 0, sizeof(*a));
  a->h = Q.n[Q.n_layers - 1];
  a->out = d->out_dim;
  a->hw = prep + Q.hw;
  a->hbias = prep + Q.hbias;
  a->log_std = prep + Q.log_std;
  a->std_x = d->std_x_coef;
  a->std_y = d->std_y_coef;
}

// advance the buffer-row-indexed pointers of a batch to chunk start c0 (identity index only)
static void head_batch(const hb_net_desc* d, const hb_actor_batch* b, int64_t c0, int64_t rows, HeadArgs* a) {
  const int ad = d->head == HB_HEAD_DISCRETE ? 1 : d->out_dim;
  const int64_t o = b->index ? 0 : c0;
  a->rows = rows;
  a->index = b->index ? b->index + c0 : nullptr;
  a->actions = b->actions ? b->actions + o * ad : nullptr;
  a->avail = b->avail ? b->avail + o * d->out_dim : nullptr;
  a->old_logp = b->old_logp ? b->old_logp + o * ad : nullptr;
  a->adv = b->adv ? b->adv + o : nullptr;
  a->factor = b->factor ? b->factor + o : nullptr;
  a->active = b->active ? b->active + o : nullptr;
}

}  // namespace hb

extern "C" {

size_t hb_workspace_bytes(const hb_net_desc* d, int64_t rows, int mode) {
  hb::PrepLayout Q;
  hb::ParamLayout P;
  if (hb::make_layouts(d, &P, &Q, nullptr)) return 0;
  int64_t ch = rows < hb::CHUNK_ROWS ? rows : hb::CHUNK_ROWS;
  if (ch < 1) ch = 1;
  return hb::work_floats(Q, ch, mode, P.total) * sizeof(float);
}

int hb_policy_act(const hb_net_desc* d, const float* prepared, const float* obs, int64_t rows, const float* avail,
                  int deterministic, uint64_t seed, uint64_t offset, float* actions, float* logp, void* ws,
                  size_t ws_bytes, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(prepared && obs && actions && logp && rows >= 0, "bad argument");
  PrepLayout Q;
  int rc = check_net(d, nullptr, &Q, nullptr, 1);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ch = rows < CHUNK_ROWS ? rows : CHUNK_ROWS;
  if (rows == 0) return HB_OK;
  Work w;
  if ((rc = carve(Q, ch, 0, ws, ws_bytes, &w))) return rc;
  const int ad = d->head == HB_HEAD_DISCRETE ? 1 : d->out_dim;
  for (int64_t c0 = 0; c0 < rows; c0 += ch) {
    const int64_t n = rows - c0 < ch ? rows - c0 : ch;
    if ((rc = trunk_forward(d, Q, prepared, obs, nullptr, c0, n, w, st))) return rc;
    HeadArgs a;
    head_base(d, Q, prepared, &a);
    a.feat = w.Y[Q.n_layers - 1];
    a.rows = n;
    a.avail = avail ? avail + c0 * d->out_dim : nullptr;
    a.deterministic = deterministic;
    a.seed = seed;
    a.offset = offset + (uint64_t)c0 * 0x9E3779B97F4A7C15ull;  // distinct Philox streams per chunk
    a.actions_out = actions + c0 * ad;
    a.logp_out = logp + c0 * ad;
    if ((rc = launch_policy_head(d->head, MODE_ACT, a, st))) return rc;
  }
  return HB_OK;
}

int hb_policy_act(const hb_net_desc* d, const float* prepared, const float* obs, int64_t rows, const float* avail,
                  int deterministic, uint64_t seed, uint64_t offset, float* actions, float* logp, void* ws,
                  size_t ws_bytes, void* stream);
int hb_value_forward(const hb_net_desc* d, const float* prepared, const float* cent_obs, int64_t rows, float* values,
                     void* ws, size_t ws_bytes, void* stream);

int hb_value_forward(const hb_net_desc* d, const float* prepared, const float* cent_obs, int64_t rows, float* values,
                     void* ws, size_t ws_bytes, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(prepared && cent_obs && values && rows >= 0, "bad argument");
  PrepLayout Q;
  int rc = check_net(d, nullptr, &Q, nullptr, 0);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) return HB_OK;
  const int64_t ch = rows < CHUNK_ROWS ? rows : CHUNK_ROWS;
  Work w;
  if ((rc = carve(Q, ch, 0, ws, ws_bytes, &w))) return rc;
  for (int64_t c0 = 0; c0 < rows; c0 += ch) {
    const int64_t n = rows - c0 < ch ? rows - c0 : ch;
    if ((rc = trunk_forward(d, Q, prepared, cent_obs, nullptr, c0, n, w, st))) return rc;
    ValueArgs a;
    memset(&a, 0, sizeof(a));
    a.feat = w.Y[Q.n_layers - 1];
    a.h = Q.n[Q.n_layers - 1];
    a.hw = prepared + Q.hw;
    a.hbias = prepared + Q.hbias;
    a.rows = n;
    a.values_out = values + c0;
    if ((rc = launch_value_head(0, a, st))) return rc;
  }
  return HB_OK;
}

int hb_policy_evaluate(const hb_net_desc* d, const float* prepared, const hb_actor_batch* b, float* logp_out,
                       const float* logp_ref, float* factor_inout, int action_aggregation_prod, void* ws,
                       size_t ws_bytes, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(prepared && b && b->obs && b->actions && b->rows >= 0, "bad argument");
  HB_CHECK_ARG(!factor_inout || (logp_ref && !b->index), "factor update needs logp_ref and an identity batch");
  PrepLayout Q;
  int rc = check_net(d, nullptr, &Q, nullptr, 1);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = b->rows;
  if (rows == 0) return HB_OK;
  const int64_t ch = rows < CHUNK_ROWS ? rows : CHUNK_ROWS;
  Work w;
  if ((rc = carve(Q, ch, 0, ws, ws_bytes, &w))) return rc;
  const int ad = d->head == HB_HEAD_DISCRETE ? 1 : d->out_dim;
  for (int64_t c0 = 0; c0 < rows; c0 += ch) {
    const int64_t n = rows - c0 < ch ? rows - c0 : ch;
    if ((rc = trunk_forward(d, Q, prepared, b->obs, b->index, c0, n, w, st))) return rc;
    HeadArgs a;
    head_base(d, Q, prepared, &a);
    head_batch(d, b, c0, n, &a);
    a.feat = w.Y[Q.n_layers - 1];
    a.logp_out = logp_out ? logp_out + c0 * ad : nullptr;
    a.logp_ref = logp_ref ? logp_ref + c0 * ad : nullptr;
    a.factor_inout = factor_inout ? factor_inout + c0 : nullptr;
    a.agg_prod = action_aggregation_prod;
    if ((rc = launch_policy_head(d->head, MODE_EVAL, a, st))) return rc;
  }
  return HB_OK;
}

int hb_ppo_actor_grad(const hb_net_desc* d, const float* params, const float* prepared, const hb_actor_batch* b,
                      const hb_ppo_hyper* h, const double* norm3, float* grad, double* scalars, void* ws,
                      size_t ws_bytes, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(params && prepared && b && h && norm3 && grad && scalars, "NULL argument");
  HB_CHECK_ARG(b->obs && b->actions && b->old_logp && b->adv && b->active && b->rows >= 0, "incomplete batch");
  ParamLayout P;
  PrepLayout Q;
  hb_net_layout L;
  int rc = check_net(d, &P, &Q, &L, 1);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t ce = cudaMemsetAsync(grad, 0, (size_t)L.total * sizeof(float), st);
  if (ce != cudaSuccess) return cuda_fail(ce, "hb_ppo_actor_grad(memset)");
  const int64_t rows = b->rows;
  if (rows == 0) return HB_OK;
  const int64_t ch = rows < CHUNK_ROWS ? rows : CHUNK_ROWS;
  Work w;
  if ((rc = carve(Q, ch, 1, ws, ws_bytes, &w, L.total))) return rc;
  {
    ce = cudaMemsetAsync(w.dwpart, 0, (size_t)tc_dw_splits() * L.total * sizeof(float), st);
    if (ce != cudaSuccess) return cuda_fail(ce, "hb_ppo_actor_grad(memset split buffer)");
  }
  for (int64_t c0 = 0; c0 < rows; c0 += ch) {
    const int64_t n = rows - c0 < ch ? rows - c0 : ch;
    if ((rc = trunk_forward(d, Q, prepared, b->obs, b->index, c0, n, w, st))) return rc;
    HeadArgs a;
    head_base(d, Q, prepared, &a);
    head_batch(d, b, c0, n, &a);
    a.feat = w.Y[Q.n_layers - 1];
    a.agg_prod = h->action_aggregation_prod;
    a.clip = h->clip_param;
    a.entropy_coef = h->entropy_coef;
    a.use_active = h->use_policy_active_masks;
    a.use_clip = h->use_clip;
    a.norm3 = norm3;
    a.dfeat = w.dA;
    a.g_hw = grad + P.hw;
    a.g_hbias = grad + P.hbias;
    a.g_log_std = d->head == HB_HEAD_BOX ? grad + P.log_std : nullptr;
    a.scalars = scalars;
    a.ln_z = w.Z[Q.n_layers - 1]; a.ln_stats = w.stats[Q.n_layers - 1]; a.ln_w = prepared + Q.lnw[Q.n_layers - 1];
    a.g_ln_w = grad + P.lnw[Q.n_layers - 1]; a.g_ln_b = grad + P.lnb[Q.n_layers - 1]; a.ln_act = d->activation;
    a.part_delta = w.dwpart - grad; a.part_stride = w.ptotal;
    if ((rc = launch_policy_head(d->head, MODE_GRAD, a, st))) return rc;
    if ((rc = trunk_backward(d, P, Q, params, prepared, grad, n, w, st))) return rc;
  }
  if ((rc = launch_dw_reduce(grad, w.dwpart, L.total, st))) return rc;
  return launch_featnorm_fold(d, params, grad, st);
}

int hb_value_grad(const hb_net_desc* d, const float* params, const float* prepared, const hb_critic_batch* b,
                  const hb_value_hyper* h, const float* vn_state, double inv_count, float* grad, double* scalars,
                  void* ws, size_t ws_bytes, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(params && prepared && b && h && grad && scalars, "NULL argument");
  HB_CHECK_ARG(b->share_obs && b->value_preds && b->returns && b->rows >= 0, "incomplete batch");
  ParamLayout P;
  PrepLayout Q;
  hb_net_layout L;
  int rc = check_net(d, &P, &Q, &L, 0);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t ce = cudaMemsetAsync(grad, 0, (size_t)L.total * sizeof(float), st);
  if (ce != cudaSuccess) return cuda_fail(ce, "hb_value_grad(memset)");
  const int64_t rows = b->rows;
  if (rows == 0) return HB_OK;
  const int64_t ch = rows < CHUNK_ROWS ? rows : CHUNK_ROWS;
  Work w;
  if ((rc = carve(Q, ch, 1, ws, ws_bytes, &w, L.total))) return rc;
  {
    ce = cudaMemsetAsync(w.dwpart, 0, (size_t)tc_dw_splits() * L.total * sizeof(float), st);
    if (ce != cudaSuccess) return cuda_fail(ce, "hb_value_grad(memset split buffer)");
  }
  for (int64_t c0 = 0; c0 < rows; c0 += ch) {
    const int64_t n = rows - c0 < ch ? rows - c0 : ch;
    if ((rc = trunk_forward(d, Q, prepared, b->share_obs, b->index, c0, n, w, st))) return rc;
    ValueArgs a;
    memset(&a, 0, sizeof(a));
    a.feat = w.Y[Q.n_layers - 1];
    a.h = Q.n[Q.n_layers - 1];
    a.hw = prepared + Q.hw;
    a.hbias = prepared + Q.hbias;
    a.rows = n;
    const int64_t o = b->index ? 0 : c0;
    a.index = b->index ? b->index + c0 : nullptr;
    a.value_preds = b->value_preds + o;
    a.returns = b->returns + o;
    a.vn_state = vn_state;
    a.clip = h->clip_param;
    a.huber_delta = h->huber_delta;
    a.coef = (float)((double)h->value_loss_coef * inv_count);
    a.use_huber = h->use_huber_loss;
    a.use_clipped = h->use_clipped_value_loss;
    a.dfeat = w.dA;
    a.g_hw = grad + P.hw;
    a.g_hbias = grad + P.hbias;
    a.scalars = scalars;
    a.ln_z = w.Z[Q.n_layers - 1]; a.ln_stats = w.stats[Q.n_layers - 1]; a.ln_w = prepared + Q.lnw[Q.n_layers - 1];
    a.g_ln_w = grad + P.lnw[Q.n_layers - 1]; a.g_ln_b = grad + P.lnb[Q.n_layers - 1]; a.ln_act = d->activation;
    a.part_delta = w.dwpart - grad; a.part_stride = w.ptotal;
    if ((rc = launch_value_head(1, a, st))) return rc;
    if ((rc = trunk_backward(d, P, Q, params, prepared, grad, n, w, st))) return rc;
  }
  if ((rc = launch_dw_reduce(grad, w.dwpart, L.total, st))) return rc;
  return launch_featnorm_fold(d, params, grad, st);
}
}
