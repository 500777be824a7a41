// Gradient finalisation (feature-norm fold), clip_grad_norm_ + Adam, rollout-insert masks.
#include "common.cuh"

namespace hb {

int prepare_launch(const hb_net_desc* d, const float* params, float* prepared, cudaStream_t st);

// The layer-0 GEMM runs on the un-affined normalised input with W' = W diag(gamma0),
// b' = b + W beta0 (hb_net_prepare).  Map gradients w.r.t. (W', b') -- what dw_accum produced
// in the W0 / b0 slots -- back to (W0, b0, gamma0, beta0):
//   dgamma0[k] = sum_n W0[n][k] G[n][k],  dbeta0[k] = sum_n W0[n][k] gb[n],
//   dW0[n][k]  = gamma0[k] G[n][k] + beta0[k] gb[n],   db0 = gb.
__global__ void featnorm_grad_fold_kernel(const float* __restrict__ params, float* __restrict__ grad, int w0, int b0,
                                          int fnw, int fnb, int N, int K) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K; k += gridDim.x * blockDim.x) {
    const float gam = params[fnw + k], bet = params[fnb + k];
    double dg = 0.0, dbt = 0.0;  // these sums cancel heavily: accumulate in fp64 (N <= 256 terms)
    for (int n = 0; n < N; ++n) {
      const float w = params[w0 + n * K + k];
      const float G = grad[w0 + n * K + k];
      const float gb = grad[b0 + n];
      dg += (double)w * (double)G;
      dbt += (double)w * (double)gb;
      grad[w0 + n * K + k] = fmaf(gam, G, bet * gb);
    }
    grad[fnw + k] = (float)dg;
    grad[fnb + k] = (float)dbt;
  }
}

int launch_featnorm_fold(const hb_net_desc* d, const float* params, float* grad, cudaStream_t st) {
  if (!d->feature_norm) return HB_OK;
  ParamLayout P;
  int rc = make_layouts(d, &P, nullptr, nullptr);
  if (rc) return rc;
  int K = d->in_dim;
  featnorm_grad_fold_kernel<<<(K + 127) / 128, 128, 0, st>>>(params, grad, P.w[0], P.b[0], P.fn_w, P.fn_b, d->hidden[0], K);
  HB_LAUNCH_DONE(st,"featnorm_grad_fold");
  return HB_OK;
}

// the same fold for any consumer of a LayerNorm affine (fused_update.cu: W' = W diag(gamma), b' = b + W beta per layer / head)
int launch_featnorm_fold_at(const float* params, float* grad, int w0, int b0, int gw, int gb, int N, int K, cudaStream_t st) {
  featnorm_grad_fold_kernel<<<(K + 127) / 128, 128, 0, st>>>(params, grad, w0, b0, gw, gb, N, K);
  HB_LAUNCH_DONE(st, "ln_affine_grad_fold");
  return HB_OK;
}

// Total L2 norm, clip coefficient, Adam (torch single-tensor form).  Every CTA computes the whole norm itself
// (<= 340 KB of gradients out of L2, the same thread-strided order in every CTA -> the identical coefficient, bit for bit the
// one-CTA result) and then updates its own slice: no grid barrier, and the update is not limited to one SM's load / store
// bandwidth (one CTA over 20k parameters took 23 us, 1 % of the C2 update phase per 20 steps).  256 threads x 32 registers:
// small enough to run on an SM that also holds a CTA of the other stream's persistent update kernel (320 x 168 registers,
// 222 KB shared memory) -- a 1024-thread CTA had to wait for that kernel to end.
__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v, int n,
                                                         float step_size, float bc2_sqrt, float b1, float b2, float eps,
                                                         float wd, float max_norm, int use_clip,
                                                         float* __restrict__ gnorm_out) {
  __shared__ double red[32];
  __shared__ float s_coef;
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { float x = g[i]; acc += (double)x * (double)x; }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = warp_sum_d(threadIdx.x < 8 ? red[threadIdx.x] : 0.0);
    if (threadIdx.x == 0) {
      float total = (float)sqrt(t);
      if (gnorm_out && blockIdx.x == 0) gnorm_out[0] = total;
      float c = 1.f;
      if (use_clip) c = fminf(max_norm / (total + 1e-6f), 1.f);
      s_coef = c;
    }
  }
  __syncthreads();
  const float coef = s_coef;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    float gi = g[i] * coef;
    float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.f - b1);
    vi = vi * b2 + (1.f - b2) * gi * gi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

// on_policy_base_runner.py:358-433
__global__ void insert_masks_kernel(hb_insert_args a) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.n_envs) return;
  const int A = a.n_agents;
  bool all_done = true;
  for (int i = 0; i < A; ++i) all_done = all_done && (a.dones[n * A + i] != 0);
  const float mk = all_done ? 0.f : 1.f;
  for (int i = 0; i < A; ++i) {
    if (a.actor_masks_next[i]) a.actor_masks_next[i][n] = mk;
    if (a.actor_active_next[i]) a.actor_active_next[i][n] = all_done ? 1.f : (a.dones[n * A + i] ? 0.f : 1.f);
  }
  if (a.rewards != nullptr && a.ep_return != nullptr) {
    float r = 0.f;
    for (int i = 0; i < A; ++i) r += a.rewards[(int64_t)n * a.reward_stride_n + (int64_t)i * a.reward_stride_a];
    float acc = a.ep_return[n] + r / (float)A;
    if (all_done) {
      if (a.done_sum) { atomicAdd(a.done_sum, (double)acc); atomicAdd(a.done_sum + 1, 1.0); }
      acc = 0.f;
    }
    a.ep_return[n] = acc;
  }
  if (a.state_type_fp) {
    for (int i = 0; i < A; ++i) {
      if (a.critic_masks_next) a.critic_masks_next[n * A + i] = mk;
      if (a.critic_bad_next) a.critic_bad_next[n * A + i] = a.bad_transition[n * A + i] ? 0.f : 1.f;
    }
  } else {
    if (a.critic_masks_next) a.critic_masks_next[n] = mk;
    if (a.critic_bad_next) a.critic_bad_next[n] = a.bad_transition[n * A] ? 0.f : 1.f;
  }
}

// zero the hidden-state rows of finished envs (on_policy_base_runner.py:358-386)
__global__ void insert_rnn_reset_kernel(hb_insert_args a) {
  const int A = a.n_agents;
  const int64_t per_env = (int64_t)a.actor_rnn_row * A + (int64_t)a.critic_rnn_row * (a.state_type_fp ? A : 1);
  const int64_t total = per_env * a.n_envs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / per_env);
    int64_t j = i % per_env;
    bool all_done = true;
    for (int q = 0; q < A; ++q) all_done = all_done && (a.dones[n * A + q] != 0);
    if (!all_done) continue;
    if (j < (int64_t)a.actor_rnn_row * A) {
      int ag = (int)(j / a.actor_rnn_row), e = (int)(j % a.actor_rnn_row);
      if (a.actor_rnn_next[ag]) a.actor_rnn_next[ag][(int64_t)n * a.actor_rnn_row + e] = 0.f;
    } else {
      j -= (int64_t)a.actor_rnn_row * A;
      if (a.critic_rnn_next) {
        int64_t rows_per_env = a.state_type_fp ? A : 1;
        a.critic_rnn_next[(int64_t)n * rows_per_env * a.critic_rnn_row + j] = 0.f;
      }
    }
  }
}

}  // namespace hb

extern "C" {

int hb_clip_adam_step(const hb_net_desc* d, float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                      float* prepared, const hb_adam_hyper* h, float* grad_norm_out, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(d && params && grad && exp_avg && exp_avg_sq && h, "NULL argument");
  HB_CHECK_ARG(h->step >= 1, "Adam step must be >= 1");
  hb_net_layout L;
  int rc = make_layouts(d, nullptr, nullptr, &L);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const double bc1 = 1.0 - pow((double)h->beta1, (double)h->step);
  const double bc2 = 1.0 - pow((double)h->beta2, (double)h->step);
  int adam_ctas = (L.total + 1023) / 1024;   // 4 elements per thread
  adam_ctas = adam_ctas < 1 ? 1 : (adam_ctas > 128 ? 128 : adam_ctas);
  clip_adam_kernel<<<adam_ctas, 256, 0, st>>>(params, grad, exp_avg, exp_avg_sq, L.total, (float)((double)h->lr / bc1),
                                       (float)sqrt(bc2), h->beta1, h->beta2, h->eps, h->weight_decay, h->max_grad_norm,
                                       h->use_max_grad_norm, grad_norm_out);
  HB_LAUNCH_DONE(st,"hb_clip_adam_step");
  if (prepared) return prepare_launch(d, params, prepared, st);
  return HB_OK;
}

int hb_rollout_insert_masks(const hb_insert_args* a, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(a && a->dones && a->bad_transition, "NULL argument");
  HB_CHECK_ARG(a->n_envs > 0 && a->n_agents > 0 && a->n_agents <= HB_MAX_AGENTS, "bad n_envs / n_agents");
  cudaStream_t st = (cudaStream_t)stream;
  insert_masks_kernel<<<(a->n_envs + 127) / 128, 128, 0, st>>>(*a);
  HB_LAUNCH_DONE(st,"hb_rollout_insert_masks");
  if (a->actor_rnn_row > 0 || a->critic_rnn_row > 0) {
    int64_t total = ((int64_t)a->actor_rnn_row * a->n_agents + (int64_t)a->critic_rnn_row * (a->state_type_fp ? a->n_agents : 1)) * a->n_envs;
    int64_t g = (total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    insert_rnn_reset_kernel<<<(unsigned)g, 256, 0, st>>>(*a);
    HB_LAUNCH_DONE(st,"hb_rollout_insert_masks(rnn reset)");
  }
  return HB_OK;
}
}
