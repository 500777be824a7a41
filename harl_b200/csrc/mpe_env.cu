// Batched MPE simple_spread: every rollout thread's world steps in one launch and writes
// its outputs straight into the rollout-buffer slots -- the env.step either side of the hot path never touches the host.
//
// Dynamics: PettingZoo 1.22 `pettingzoo.mpe.simple_spread_v2` (the package the reference wraps,
// harl/envs/pettingzoo_mpe/pettingzoo_mpe_env.py:26-33; it is not vendored in the reference and not installed here, so
// the world model is restated from the MPE particle-world definition):
//   world  : dt 0.1, damping 0.25, contact force 1e2, contact margin 1e-3; agents size 0.15, mass 1, accel 5 (sensitivity);
//            landmarks do not collide.
//   action : Discrete(5) -> u = (0,0), (-1,0), (+1,0), (0,-1), (0,+1); continuous (5) -> u = (a1 - a2, a3 - a4); u *= 5.
//   step   : collision forces between agents  f = 1e2 * d/|d| * softplus_k(0.3 - |d|), k = 1e-3;
//            v = v * 0.75 + (u + f) * 0.1;  p += v * 0.1
//   obs_i  : [v_i, p_i, landmarks - p_i, other agents - p_i, other agents' comm (zeros)]   (4 + 2L + 4(A-1) floats)
//   reward : per agent 0.5 * global + 0.5 * local,  global = -sum_l min_a |p_a - l|,  local_i = -#collisions of agent i.
// Adapter semantics of the reference wrapper (pettingzoo_mpe_env.py:41-88) and its vector env
// (harl/envs/env_wrappers.py: auto-reset on done): team reward = sum over agents broadcast to every agent, truncation after
// `max_cycles` (25) steps with bad_transition, state = concatenation of all agents' observations repeated per agent,
// observations returned for a finished env are those of the freshly reset world.
// One thread per world; positions from Philox4x32-10 keyed by (seed, world, episode): the NumPy twin
// (harl_b200/envs/mpe_spread.py) draws the same numbers.
#include "common.cuh"

namespace hb {

constexpr int MPE_MAX_A = 8;

__device__ __forceinline__ double mpe_u(uint32_t x) { return 2.0 * (double)u01(x) - 1.0; }

// initial positions of episode `ep` of world `n`: agents then landmarks, (x, y) pairs, 4 uniforms per Philox call
__device__ __forceinline__ void mpe_reset_world(const hb_mpe_args& a, int64_t n, uint64_t ep, float* pos, float* vel, float* lm) {
  const int A = a.n_agents, Lm = a.n_landmarks;
  const int pairs = A + Lm;
  for (int q = 0; q < (pairs + 1) / 2; ++q) {
    const uint4 r = philox4x32(make_uint4((uint32_t)n, (uint32_t)((uint64_t)n >> 32), (uint32_t)q, (uint32_t)ep),
                               make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ (uint32_t)(ep >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    for (int h = 0; h < 2; ++h) {
      const int e = 2 * q + h;
      if (e >= pairs) break;
      float* dst = e < A ? pos + 2 * e : lm + 2 * (e - A);
      dst[0] = (float)mpe_u(w[2 * h]);
      dst[1] = (float)mpe_u(w[2 * h + 1]);
    }
  }
  for (int i = 0; i < 2 * A; ++i) vel[i] = 0.f;
}

__global__ void mpe_spread_kernel(hb_mpe_args a) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.n_envs) return;
  const int A = a.n_agents, Lm = a.n_landmarks;
  float pos[2 * MPE_MAX_A], vel[2 * MPE_MAX_A], lm[2 * MPE_MAX_A];
  float* gp = a.pos + n * 2 * A;
  float* gv = a.vel + n * 2 * A;
  float* gl = a.landmarks + n * 2 * Lm;
  float team = 0.f;
  bool done = false;
  if (a.reset_all) {
    a.episode[n] = 0;
    a.step_count[n] = 0;
    mpe_reset_world(a, n, 0, pos, vel, lm);
  } else {
    for (int i = 0; i < 2 * A; ++i) { pos[i] = gp[i]; vel[i] = gv[i]; }
    for (int i = 0; i < 2 * Lm; ++i) lm[i] = gl[i];
    // ---- actions -> control forces
    double fx[MPE_MAX_A], fy[MPE_MAX_A];
    for (int i = 0; i < A; ++i) {
      double ux = 0.0, uy = 0.0;
      if (a.continuous) {
        const float* ac = a.actions[i] + n * 5;
        ux = (double)ac[1] - (double)ac[2];
        uy = (double)ac[3] - (double)ac[4];
      } else {
        const int k = (int)a.actions[i][n];
        ux = k == 1 ? -1.0 : (k == 2 ? 1.0 : 0.0);
        uy = k == 3 ? -1.0 : (k == 4 ? 1.0 : 0.0);
      }
      fx[i] = 5.0 * ux;
      fy[i] = 5.0 * uy;
    }
    // ---- collision forces between agents (core.py get_collision_force)
    for (int i = 0; i < A; ++i)
      for (int j = i + 1; j < A; ++j) {
        const double dx = (double)pos[2 * i] - (double)pos[2 * j], dy = (double)pos[2 * i + 1] - (double)pos[2 * j + 1];
        const double dist = sqrt(dx * dx + dy * dy);
        const double k = 1e-3, x = -(dist - 0.3) / k;
        const double pen = (x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x))) * k;   // logaddexp(0, x) * k
        const double f = 1e2 * pen / dist;
        fx[i] += f * dx; fy[i] += f * dy;
        fx[j] -= f * dx; fy[j] -= f * dy;
      }
    // ---- integrate
    for (int i = 0; i < A; ++i) {
      const double vx = (double)vel[2 * i] * 0.75 + fx[i] * 0.1, vy = (double)vel[2 * i + 1] * 0.75 + fy[i] * 0.1;
      vel[2 * i] = (float)vx; vel[2 * i + 1] = (float)vy;
      pos[2 * i] = (float)((double)pos[2 * i] + vx * 0.1);
      pos[2 * i + 1] = (float)((double)pos[2 * i + 1] + vy * 0.1);
    }
    // ---- rewards on the new state
    double glob = 0.0;
    for (int l = 0; l < Lm; ++l) {
      double best = 1e30;
      for (int i = 0; i < A; ++i) {
        const double dx = (double)pos[2 * i] - (double)lm[2 * l], dy = (double)pos[2 * i + 1] - (double)lm[2 * l + 1];
        best = fmin(best, sqrt(dx * dx + dy * dy));
      }
      glob -= best;
    }
    double total = 0.0;
    for (int i = 0; i < A; ++i) {
      double local = 0.0;
      for (int j = 0; j < A; ++j) {
        if (j == i) continue;
        const double dx = (double)pos[2 * i] - (double)pos[2 * j], dy = (double)pos[2 * i + 1] - (double)pos[2 * j + 1];
        if (sqrt(dx * dx + dy * dy) < 0.3) local -= 1.0;
      }
      const double r = 0.5 * glob + 0.5 * local;
      total += r;
    }
    team = (float)total;
    const int s = a.step_count[n] + 1;
    done = s >= a.max_cycles;
    if (done) {  // truncation: the vector env resets the world and returns the new episode's observations
      const uint64_t ep = a.episode[n] + 1;
      a.episode[n] = ep;
      a.step_count[n] = 0;
      mpe_reset_world(a, n, ep, pos, vel, lm);
    } else {
      a.step_count[n] = s;
    }
  }
  for (int i = 0; i < 2 * A; ++i) { gp[i] = pos[i]; gv[i] = vel[i]; }
  for (int i = 0; i < 2 * Lm; ++i) gl[i] = lm[i];
  // ---- observations / state
  const int od = 4 + 2 * Lm + 4 * (A - 1);
  for (int i = 0; i < A; ++i) {
    float* o = a.obs_out[i] + n * od;
    float* so = a.share_obs_out ? a.share_obs_out + n * (int64_t)(A * od) + i * od : nullptr;
    int c = 0;
    auto put = [&](float v) { o[c] = v; if (so) so[c] = v; ++c; };
    put(vel[2 * i]); put(vel[2 * i + 1]);
    put(pos[2 * i]); put(pos[2 * i + 1]);
    for (int l = 0; l < Lm; ++l) { put(lm[2 * l] - pos[2 * i]); put(lm[2 * l + 1] - pos[2 * i + 1]); }
    for (int j = 0; j < A; ++j) if (j != i) { put(pos[2 * j] - pos[2 * i]); put(pos[2 * j + 1] - pos[2 * i + 1]); }
    for (int j = 0; j < 2 * (A - 1); ++j) put(0.f);
  }
  if (!a.reset_all) {
    if (a.rewards_out) a.rewards_out[n] = team;
    for (int i = 0; i < A; ++i) {
      if (a.rewards_na_out) a.rewards_na_out[n * A + i] = team;
      if (a.dones_out) a.dones_out[n * A + i] = done ? 1 : 0;
      if (a.bad_out) a.bad_out[n * A + i] = done ? 1 : 0;
    }
  }
}

}  // namespace hb

extern "C" int hb_mpe_spread_step(const hb_mpe_args* a, void* stream) {
  using namespace hb;
  HB_CHECK_ARG(a && a->pos && a->vel && a->landmarks && a->step_count && a->episode, "NULL state");
  HB_CHECK_ARG(a->n_agents >= 1 && a->n_agents <= MPE_MAX_A && a->n_landmarks >= 1 && a->n_landmarks <= MPE_MAX_A, "1..8 agents / landmarks");
  HB_CHECK_ARG(a->n_envs >= 0 && a->max_cycles >= 1, "bad sizes");
  for (int i = 0; i < a->n_agents; ++i) HB_CHECK_ARG(a->obs_out[i] && (a->reset_all || a->actions[i]), "NULL per-agent pointer");
  if (a->n_envs == 0) return HB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  mpe_spread_kernel<<<(unsigned)((a->n_envs + 127) / 128), 128, 0, st>>>(*a);
  HB_LAUNCH_DONE(st, "mpe_spread_step");
  return HB_OK;
}
