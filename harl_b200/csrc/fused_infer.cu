// Fused rollout inference: ONE launch computes, for every agent's actor and for the critic,
//   feature LayerNorm -> [Linear -> act -> LayerNorm] x L -> head (sample / mode + log-prob, or value)
// for all rollout threads (OnPolicyBaseRunner.collect, harl/runners/on_policy_base_runner.py:285-340, which issues
// (A+1) x ~10 framework kernels per step).  Rollout batches are small (N rows per net), so the step is launch- and
// latency-bound: grid = (row tiles of 64) x (A+1 nets); activations never leave shared memory; weights (<= 340 KB per
// net, L2-resident) stream through a double-buffered shared tile.  FP32 FFMA with the same summation order as
// linear_ln_fwd_kernel / the row-wise head kernels, so results are bit-identical to the unfused path.
#include <math.h>

#include "common.cuh"
#include "head_rows.cuh"
#include "fused_args.cuh"
#include "kernels.cuh"

namespace hb {

constexpr int FI_ROWS = 64;             // rows per CTA: 4 per thread -> 6.4 FMAs per shared-memory wavefront (FFMA-bound)
constexpr int FI_RPT = FI_ROWS / 16;   // rows per thread
constexpr int FI_KC = 16;
#define FI_LOG_2PI_F 1.8378770664093453f

struct InferNet {
  const float* prep;
  const float* obs;      // [rows, in_dim]
  const float* avail;    // [rows, out] or null
  float* out0;           // actions [rows, ad] | values [rows]
  float* out1;           // log-probs [rows, ad] | unused
  unsigned long long seed;
  long long rows;
  int in_dim, out_dim, head;
  float std_x, std_y;
};

struct InferArgs {
  int n_nets, n_layers, act, feature_norm, deterministic;
  int hidden[HB_MAX_LAYERS];
  unsigned long long offset;
  const unsigned long long* offset_base;   // device counter added to offset (nullable)
  InferNet net[HB_MAX_AGENTS + 1];
};

// Categorical head on a row group (16 lanes per row, 2 rows per warp): the same arithmetic as discrete_rows_kernel's
// act mode (head_rows.cuh), fed from the activation tile in shared memory.
template <int CPL>
__device__ __forceinline__ void fused_discrete_rows(const InferArgs& A, const InferNet& net, const float* __restrict__ cur,
                                                    int pc, const float* __restrict__ shw, const float* __restrict__ sb,
                                                    int nrows, long long r0, int od) {
  constexpr int LPR = 16, MAXJ = 8, NC = CPL / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = lane % LPR, rw = lane / LPR;
  const unsigned long long off = A.offset + (A.offset_base ? *A.offset_base : 0ull);
  for (int rr = warp * 2; rr < FI_ROWS; rr += 16) {
    const int r = rr + rw;
    const bool ok = r < nrows;
    const long long row = r0 + r;
    float f[CPL];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 v = rows::ld4(cur + r * pc + c * 4 * LPR + 4 * s);
      f[c * 4 + 0] = v.x; f[c * 4 + 1] = v.y; f[c * 4 + 2] = v.z; f[c * 4 + 3] = v.w;
    }
    float lg[MAXJ];
    rows::group_dots<CPL, LPR, MAXJ>(f, shw, s, lg);
    float av = 1.f;
    if (ok && net.avail != nullptr && s < od) av = net.avail[row * od + s];
    const unsigned avm = net.avail != nullptr ? ((__ballot_sync(rows::FULL, av != 0.f) >> (rw * LPR)) & 0xffu) : 0xffu;
    float lp[MAXJ], pj[MAXJ];
    rows::categorical<MAXJ>(lg, sb, od, avm, lp, pj);
    const int pick = rows::categorical_pick<MAXJ>(pj, od, A.deterministic != 0,
                                                  A.deterministic ? 0.f : rows::row_uniform(row, net.seed, off));
    const float lpp = rows::select<MAXJ>(lp, pick);
    if (ok && s == 0) { net.out0[row] = (float)pick; net.out1[row] = lpp; }
  }
}

template <int NT>
__global__ void __launch_bounds__(256) fused_infer_kernel(const __grid_constant__ InferArgs A) {
  extern __shared__ __align__(16) float fsm[];
  const InferNet& net = A.net[blockIdx.y];
  const long long r0 = (long long)blockIdx.x * FI_ROWS;
  if (r0 >= net.rows) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int nrows = (int)(net.rows - r0 < FI_ROWS ? net.rows - r0 : FI_ROWS);
  const int in_dim = net.in_dim, kp0 = (in_dim + 3) & ~3;
  const int pin = kp0 + 4;                  // input tile pitch (floats)
  constexpr int PA = NT + 4;                // activation tile pitch
  float* xin = fsm;                         // [32][pin]
  float* actA = xin + FI_ROWS * pin;        // [32][PA]
  float* actB = actA + FI_ROWS * PA;        // [32][PA]
  float* wbuf = actB + FI_ROWS * PA;        // [2][FI_KC][NT]
  constexpr int NCH = NT / 64;

  // ---- observations -> shared, feature LayerNorm (affine folded into layer 0 by hb_net_prepare)
  {
    const float* src = net.obs + r0 * in_dim;
    for (int f = tid; f < nrows * in_dim; f += 256) xin[(f / in_dim) * pin + f % in_dim] = src[f];
    for (int f = tid; f < FI_ROWS * pin; f += 256) {
      int r = f / pin, k = f % pin;
      if (r >= nrows || k >= in_dim) xin[f] = 0.f;
    }
    __syncthreads();
    if (A.feature_norm) {
      if (in_dim <= 44) {  // same arithmetic as feat_norm_narrow_kernel (thread per row)
        if (tid < nrows) {
          float* o = xin + tid * pin;
          float s = 0.f;
          for (int k = 0; k < in_dim; ++k) s += o[k];
          const float mean = s / (float)in_dim;
          float q = 0.f;
          for (int k = 0; k < in_dim; ++k) { float d = o[k] - mean; q = fmaf(d, d, q); }
          const float rstd = rsqrtf(q / (float)in_dim + 1e-5f);
          for (int k = 0; k < in_dim; ++k) o[k] = (o[k] - mean) * rstd;
        }
      } else {             // same arithmetic as feat_norm_kernel (warp per row)
        for (int r = warp; r < nrows; r += 8) {
          float* o = xin + r * pin;
          float s = 0.f;
          for (int k = lane; k < in_dim; k += 32) s += o[k];
          const float mean = warp_sum(s) / (float)in_dim;
          float q = 0.f;
          for (int k = lane; k < in_dim; k += 32) { float d = o[k] - mean; q = fmaf(d, d, q); }
          const float rstd = rsqrtf(warp_sum(q) / (float)in_dim + 1e-5f);
          for (int k = lane; k < in_dim; k += 32) o[k] = (o[k] - mean) * rstd;
        }
      }
    }
    __syncthreads();
  }

  // ---- trunk
  const float* cur = xin;
  int pc = pin, K = kp0, off = 0;
  float* nxt = actA;
  for (int l = 0; l < A.n_layers; ++l) {
    const int N = A.hidden[l];
    const float* WT = net.prep + off;           // [K][N]
    const float* bias = WT + (size_t)K * N;
    const float* lnw = bias + N;
    const float* lnb = lnw + N;
    off += K * N + 3 * N;
    float acc[FI_RPT][NT / 16];
#pragma unroll
    for (int i = 0; i < FI_RPT; ++i)
#pragma unroll
      for (int j = 0; j < NT / 16; ++j) acc[i][j] = 0.f;
    const int nk = (K + FI_KC - 1) / FI_KC;
    constexpr int BLD = FI_KC * NT / 4 / 256;   // float4 of the weight tile per thread
    float4 rb[BLD];
    auto load_b = [&](int k0) {
#pragma unroll
      for (int q = 0; q < BLD; ++q) {
        const int f = tid + q * 256;
        const int kk = f / (NT / 4), n4 = (f % (NT / 4)) * 4;
        rb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + kk < K && n4 < N) rb[q] = *reinterpret_cast<const float4*>(WT + (size_t)(k0 + kk) * N + n4);
      }
    };
    auto store_b = [&](int buf) {
#pragma unroll
      for (int q = 0; q < BLD; ++q) {
        const int f = tid + q * 256;
        const int kk = f / (NT / 4), n4 = (f % (NT / 4)) * 4;
        *reinterpret_cast<float4*>(&wbuf[(buf * FI_KC + kk) * NT + n4]) = rb[q];
      }
    };
    load_b(0);
    store_b(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
      const int buf = kt & 1;
      if (kt + 1 < nk) load_b((kt + 1) * FI_KC);
#pragma unroll
      for (int k4 = 0; k4 < FI_KC; k4 += 4) {
        const int kg = kt * FI_KC + k4;
        float4 av[FI_RPT];
#pragma unroll
        for (int i = 0; i < FI_RPT; ++i)
          av[i] = kg < K ? *reinterpret_cast<const float4*>(&cur[(ty * FI_RPT + i) * pc + kg]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float4 bv[NCH];
#pragma unroll
          for (int c = 0; c < NCH; ++c) bv[c] = *reinterpret_cast<const float4*>(&wbuf[(buf * FI_KC + k4 + kk) * NT + c * 64 + tx * 4]);
#pragma unroll
          for (int i = 0; i < FI_RPT; ++i) {
            const float a = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
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
      if (kt + 1 < nk) store_b(buf ^ 1);
      __syncthreads();
    }
    // epilogue: bias, activation, LayerNorm (row statistics inside a half-warp), write to the other tile
    const float inv_n = 1.f / (float)N;
#pragma unroll
    for (int i = 0; i < FI_RPT; ++i) {
      const int r = ty * FI_RPT + i;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int n = c * 64 + tx * 4;
        const bool ok = n < N;
        float4 b4 = ok ? *reinterpret_cast<const float4*>(bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[i][c * 4 + 0] = ok ? act_fwd_rt(A.act, acc[i][c * 4 + 0] + b4.x) : 0.f;
        acc[i][c * 4 + 1] = ok ? act_fwd_rt(A.act, acc[i][c * 4 + 1] + b4.y) : 0.f;
        acc[i][c * 4 + 2] = ok ? act_fwd_rt(A.act, acc[i][c * 4 + 2] + b4.z) : 0.f;
        acc[i][c * 4 + 3] = ok ? act_fwd_rt(A.act, acc[i][c * 4 + 3] + b4.w) : 0.f;
        sum += acc[i][c * 4 + 0] + acc[i][c * 4 + 1] + acc[i][c * 4 + 2] + acc[i][c * 4 + 3];
      }
      const float mean = half_warp_sum(sum) * inv_n;
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (c * 64 + tx * 4 < N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { float d = acc[i][c * 4 + j] - mean; sq = fmaf(d, d, sq); }
        }
      const float rstd = rsqrtf(half_warp_sum(sq) * inv_n + 1e-5f);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int n = c * 64 + tx * 4;
        if (n < N) {
          const float4 g = *reinterpret_cast<const float4*>(lnw + n), be = *reinterpret_cast<const float4*>(lnb + n);
          float4 y;
          y.x = (acc[i][c * 4 + 0] - mean) * rstd * g.x + be.x;
          y.y = (acc[i][c * 4 + 1] - mean) * rstd * g.y + be.y;
          y.z = (acc[i][c * 4 + 2] - mean) * rstd * g.z + be.z;
          y.w = (acc[i][c * 4 + 3] - mean) * rstd * g.w + be.w;
          *reinterpret_cast<float4*>(&nxt[r * PA + n]) = y;
        }
      }
    }
    __syncthreads();
    cur = nxt;
    pc = PA;
    K = N;
    nxt = (nxt == actA) ? actB : actA;
  }

  // ---- head (warp per row; same arithmetic as the *_head_kernel ACT / forward modes)
  const int h = A.hidden[A.n_layers - 1];
  const int od = net.out_dim;
  const float* hw = net.prep + off;
  const float* hbias = hw + ((od * h + 3) & ~3);
  const float* log_std = hbias + ((od + 3) & ~3);
  if (net.head == HB_HEAD_DISCRETE && od <= 8 && (h == 128 || h == 64) && h <= NT) {
    float* shw = wbuf;             // [8][h] zero padded (the weight ring is idle now), then the 8 biases
    float* sb = wbuf + 8 * h;
    for (int i = tid; i < 8 * h; i += 256) shw[i] = i < od * h ? hw[i] : 0.f;
    if (tid < 8) sb[tid] = tid < od ? hbias[tid] : 0.f;
    __syncthreads();
    if (h == 128) fused_discrete_rows<8>(A, net, cur, pc, shw, sb, nrows, r0, od);
    else fused_discrete_rows<4>(A, net, cur, pc, shw, sb, nrows, r0, od);
    return;
  }
  for (int r = warp; r < nrows; r += 8) {
    const long long row = r0 + r;
    const float* f = cur + r * pc;
    float mine = 0.f;
    for (int j = 0; j < od; ++j) {
      float p = 0.f;
      for (int n = lane; n < h; n += 32) p = fmaf(f[n], hw[j * h + n], p);
      p = warp_sum(p);
      if (lane == j) mine = p + hbias[j];
    }
    const bool valid = lane < od;
    if (net.head == HB_HEAD_VALUE) {
      if (lane == 0) net.out0[row] = mine;
    } else if (net.head == HB_HEAD_DISCRETE) {
      float logit = mine;
      if (valid && net.avail != nullptr && net.avail[row * od + lane] == 0.f) logit = -1e10f;
      const float mx = warp_max(valid ? logit : -INFINITY);
      const float ex = valid ? expf(logit - mx) : 0.f;
      const float lse = mx + logf(warp_sum(ex));
      const float lp = valid ? logit - lse : 0.f;
      const float p = valid ? expf(lp) : 0.f;
      int act;
      if (A.deterministic) {
        const float pm = warp_max(p);
        act = __ffs(__ballot_sync(0xffffffffu, valid && p == pm)) - 1;
      } else {
        const unsigned long long off = A.offset + (A.offset_base ? *A.offset_base : 0ull);
        const uint4 rnd = philox4x32(make_uint4((uint32_t)row, (uint32_t)((unsigned long long)row >> 32), 0u, (uint32_t)off),
                                     make_uint2((uint32_t)net.seed, (uint32_t)(net.seed >> 32) ^ (uint32_t)(off >> 32)));
        const float u = u01(rnd.x);
        float c = p;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, c, o); if (lane >= o) c += t; }
        const unsigned below = __ballot_sync(0xffffffffu, valid && c < u);
        const unsigned pos = __ballot_sync(0xffffffffu, valid && p > 0.f);
        const int last = 31 - __clz(pos);
        act = __popc(below);
        if (act > last) act = last;
        while (act < 31 && !((pos >> act) & 1u)) ++act;
      }
      const float lpa = __shfl_sync(0xffffffffu, lp, act);
      if (lane == 0) { net.out0[row] = (float)act; net.out1[row] = lpa; }
    } else {  // DiagGaussian
      float sig = 0.f, std = 1.f;
      if (valid) { sig = 1.f / (1.f + expf(-log_std[lane] / net.std_x)); std = sig * net.std_y; }
      const float log_std_v = logf(std);
      float act = mine;
      if (!A.deterministic) {
        const unsigned long long off = A.offset + (A.offset_base ? *A.offset_base : 0ull);
        const uint4 rnd = philox4x32(make_uint4((uint32_t)row, (uint32_t)((unsigned long long)row >> 32), (uint32_t)lane, (uint32_t)off),
                                     make_uint2((uint32_t)net.seed, (uint32_t)(net.seed >> 32) ^ (uint32_t)(off >> 32)));
        const float u1 = u01(rnd.x), u2 = u01(rnd.y);
        act = mine + std * (sqrtf(-2.f * logf(u1)) * cospif(2.f * u2));
      }
      if (valid) {
        const float d = act - mine;
        net.out0[row * od + lane] = act;
        net.out1[row * od + lane] = -(d * d) / (2.f * std * std) - log_std_v - 0.5f * FI_LOG_2PI_F;
      }
    }
  }
}

static size_t fused_infer_smem(int nt, int kp0_max) {
  return sizeof(float) * ((size_t)FI_ROWS * (kp0_max + 4) + 2 * (size_t)FI_ROWS * (nt + 4) + 2 * (size_t)FI_KC * nt);
}

}  // namespace hb

namespace hb {
__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) { *c += inc; }
}  // namespace hb

extern "C" int hb_counter_add(uint64_t* counter, uint64_t inc, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(counter != nullptr, "NULL counter");
  cudaStream_t st = (cudaStream_t)stream;
  counter_add_kernel<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(counter), (unsigned long long)inc);
  HB_LAUNCH_DONE(st, "hb_counter_add");
  return HB_OK;
}

extern "C" int hb_rollout_collect(const hb_collect_args* a, void* ws, size_t ws_bytes, void* stream) {
  using namespace hb;
  (void)ws; (void)ws_bytes;
  HB_CHECK_ARG(a && a->n_agents > 0 && a->n_agents <= HB_MAX_AGENTS && a->rows > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  static thread_local InferArgs A;  // large POD: avoid re-zeroing 3 KB of stack per step
  const hb_net_desc* d0 = a->actor_desc[0];
  HB_CHECK_ARG(d0 != nullptr, "actor_desc[0] is NULL");
  bool any_rnn = a->critic_desc != nullptr && a->critic_desc->rnn_layers != 0;
  for (int i = 0; i < a->n_agents; ++i) any_rnn = any_rnn || (a->actor_desc[i] != nullptr && a->actor_desc[i]->rnn_layers != 0);
  if (any_rnn) {
    // Recurrent nets: the GRU cell is not part of the fused kernel yet -- one library call still covers all agents
    // and the critic (no host work between them), through the per-net kernels.
    for (int i = 0; i < a->n_agents; ++i) {
      int rc = hb_policy_act_rnn(a->actor_desc[i], a->actor_prepared[i], a->obs[i], a->rows, a->avail[i], a->actor_rnn[i],
                                 a->actor_masks[i], a->deterministic, a->seed[i], a->offset, a->offset_base, a->actions[i],
                                 a->logp[i], a->actor_rnn_out[i], ws, ws_bytes, stream);
      if (rc) return rc;
    }
    if (a->critic_desc != nullptr)
      return hb_value_forward_rnn(a->critic_desc, a->critic_prepared, a->share_obs, a->critic_rows, a->critic_rnn,
                                  a->critic_masks, a->values, a->critic_rnn_out, ws, ws_bytes, stream);
    return HB_OK;
  }
  // tensor-core path: every net inside the fused kernel's shapes, one trunk width / activation for all of them
  if (fused_enabled() && a->n_agents + (a->critic_desc ? 1 : 0) <= fz::ACT_MAX_NETS) {
    static thread_local fz::ActArgs F;
    F.n_nets = a->n_agents + (a->critic_desc ? 1 : 0);
    F.H = d0->hidden[0]; F.act = d0->activation; F.deterministic = a->deterministic;
    F.offset = a->offset; F.offset_base = reinterpret_cast<const unsigned long long*>(a->offset_base);
    bool all = true;
    for (int i = 0; all && i < F.n_nets; ++i) {
      const bool critic = i == a->n_agents;
      const hb_net_desc* d = critic ? a->critic_desc : a->actor_desc[i];
      bool ok = false;
      int rc = fused_act_fill(&F.net[i], d, critic ? a->critic_prepared : a->actor_prepared[i], critic ? a->share_obs : a->obs[i],
                              critic ? nullptr : a->avail[i], critic ? a->values : a->actions[i], critic ? nullptr : a->logp[i],
                              critic ? 0ull : a->seed[i], critic ? a->critic_rows : a->rows, &ok);
      if (rc) return rc;
      all = ok && d->hidden[0] == F.H && d->activation == F.act;
      if (all) HB_CHECK_ARG(F.net[i].prep && F.net[i].obs && F.net[i].out0 && (critic || F.net[i].out1), "NULL buffer");
    }
    if (all) return launch_fused_act(F, st);
  }
  A.n_nets = a->n_agents + (a->critic_desc ? 1 : 0);
  A.n_layers = d0->n_layers;
  A.act = d0->activation;
  A.feature_norm = d0->feature_norm;
  A.deterministic = a->deterministic;
  A.offset = a->offset;
  A.offset_base = reinterpret_cast<const unsigned long long*>(a->offset_base);
  int hmax = 0, kp0 = 0;
  long long max_rows = a->rows;
  for (int l = 0; l < d0->n_layers; ++l) { A.hidden[l] = d0->hidden[l]; hmax = d0->hidden[l] > hmax ? d0->hidden[l] : hmax; }
  for (int i = 0; i < A.n_nets; ++i) {
    const bool critic = i == a->n_agents;
    const hb_net_desc* d = critic ? a->critic_desc : a->actor_desc[i];
    PrepLayout Q;
    int rc = make_layouts(d, nullptr, &Q, nullptr);
    if (rc) return rc;
    bool same = d->n_layers == d0->n_layers && d->activation == d0->activation && d->feature_norm == d0->feature_norm;
    for (int l = 0; same && l < d->n_layers; ++l) same = d->hidden[l] == d0->hidden[l];
    if (!same) { set_error("fused rollout inference needs one trunk architecture for all actors and the critic"); return HB_ERR_UNSUPPORTED; }
    InferNet& n = A.net[i];
    n.prep = critic ? a->critic_prepared : a->actor_prepared[i];
    n.obs = critic ? a->share_obs : a->obs[i];
    n.avail = critic ? nullptr : a->avail[i];
    n.out0 = critic ? a->values : a->actions[i];
    n.out1 = critic ? nullptr : a->logp[i];
    n.seed = critic ? 0ull : a->seed[i];
    n.rows = critic ? a->critic_rows : a->rows;
    n.in_dim = d->in_dim;
    n.out_dim = d->out_dim;
    n.head = d->head;
    n.std_x = d->std_x_coef;
    n.std_y = d->std_y_coef;
    HB_CHECK_ARG(n.prep && n.obs && n.out0 && (critic || n.out1), "NULL buffer");
    if (Q.kpad[0] > kp0) kp0 = Q.kpad[0];
    if (n.rows > max_rows) max_rows = n.rows;
  }
  const int nt = hmax <= 64 ? 64 : hmax <= 128 ? 128 : 256;
  const size_t smem = fused_infer_smem(nt, kp0);
  if (smem > 220 * 1024) { set_error("fused rollout inference: input width %d does not fit shared memory", kp0); return HB_ERR_UNSUPPORTED; }
  dim3 grid((unsigned)((max_rows + FI_ROWS - 1) / FI_ROWS), (unsigned)A.n_nets);
#define HB_FI(NTV)                                                                                          \
  case NTV: {                                                                                               \
    static bool attr_done = false;                                                                          \
    if (!attr_done) {                                                                   \
      cudaFuncSetAttribute(fused_infer_kernel<NTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); \
      attr_done = true;                                                                                     \
    }                                                                                                       \
    fused_infer_kernel<NTV><<<grid, 256, smem, st>>>(A);                                                    \
  } break;
  switch (nt) { HB_FI(64) HB_FI(128) HB_FI(256) }
#undef HB_FI
  HB_LAUNCH_DONE(st, "fused_infer");
  return HB_OK;
}
