"""Batched MPE ``simple_spread`` behind the reference's env contract.

Two implementations of the same world model (stated in ``harl_b200/csrc/mpe_env.cu``; adapter semantics of
harl/envs/pettingzoo_mpe/pettingzoo_mpe_env.py:41-88 and the auto-reset of harl/envs/env_wrappers.py):

* ``BatchedSimpleSpread`` -- all rollout threads step in ONE kernel launch (``hb_mpe_spread_step``); with ``step_into`` the
  observations / state / team reward / done flags land directly in the rollout-buffer slots, so a whole rollout is
  CUDA-graph capturable (``graph_period() == 1``).  This is what ``--env pettingzoo_mpe`` trains on when
  ``env_args['backend'] == 'native'``.
* ``SimpleSpreadNumpy`` -- the NumPy twin with the reference's host-side call contract (lists of dict infos), used to run
  the UNMODIFIED reference on the same task (baseline/ref_runner.py) and to pin the kernel
  (tests/test_mpe_spread.py: identical seeds -> identical episodes).

Both draw the initial positions from Philox4x32-10 keyed by (seed, env index, episode), so a device env and its twin
generate the same worlds.
"""
import ctypes as C

import numpy as np

from .spaces import Box, Discrete

MAX_CYCLES = 25


def obs_dim(n_agents, n_landmarks):
    return 4 + 2 * n_landmarks + 4 * (n_agents - 1)


# ------------------------------------------------------------------ Philox4x32-10 (same stream as common.cuh)
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32(ctr, key):
    """ctr: [..., 4] uint32 array, key: [..., 2] uint32 -> [..., 4] uint32 (10 rounds)."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k = [key[..., i].astype(np.uint64) for i in range(2)]
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_M0) * c[0]
        p1 = np.uint64(_M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + np.uint64(_W0)) & mask, (k[1] + np.uint64(_W1)) & mask]
    return np.stack(c, axis=-1).astype(np.uint32)


def initial_positions(seed, env_ids, episodes, n_agents, n_landmarks):
    """Agents' and landmarks' start positions of the given (env, episode) pairs: ([n, A, 2], [n, L, 2]) float32."""
    env_ids = np.asarray(env_ids, np.uint64)
    episodes = np.asarray(episodes, np.uint64)
    pairs = n_agents + n_landmarks
    out = np.zeros((env_ids.shape[0], pairs, 2), np.float32)
    key = np.stack([np.full_like(env_ids, seed & 0xFFFFFFFF), ((seed >> 32) & 0xFFFFFFFF) ^ (episodes >> np.uint64(32))], -1)
    for q in range((pairs + 1) // 2):
        ctr = np.stack([env_ids & np.uint64(0xFFFFFFFF), env_ids >> np.uint64(32), np.full_like(env_ids, q),
                        episodes & np.uint64(0xFFFFFFFF)], -1)
        r = philox4x32(ctr.astype(np.uint32), key.astype(np.uint32))
        u = ((r >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)   # (0, 1], common.cuh u01
        v = (2.0 * u.astype(np.float64) - 1.0).astype(np.float32)
        for h in range(2):
            e = 2 * q + h
            if e < pairs:
                out[:, e, 0], out[:, e, 1] = v[:, 2 * h], v[:, 2 * h + 1]
    return out[:, :n_agents].copy(), out[:, n_agents:].copy()


class SimpleSpreadNumpy:
    """Host twin: the reference's batched-env contract (envs_tools.py:51-54), NumPy in / NumPy out."""

    team_reward = True

    def __init__(self, seed, n_threads, env_args=None, env_offset=0):
        env_args = env_args or {}
        self.n_agents = A = int(env_args.get("n_agents", 3))
        self.n_landmarks = L = int(env_args.get("n_landmarks", A))
        self.N = self.n_threads = int(n_threads)
        self.discrete = not bool(env_args.get("continuous_actions", False))
        self.max_cycles = int(env_args.get("max_cycles", MAX_CYCLES))
        self.seed_value = int(seed)
        self.env_ids = np.arange(env_offset, env_offset + self.N)
        od = obs_dim(A, L)
        self.observation_space = [Box(shape=(od,)) for _ in range(A)]
        self.share_observation_space = [Box(shape=(A * od,)) for _ in range(A)]
        self.action_space = [Discrete(5) if self.discrete else Box(shape=(5,)) for _ in range(A)]
        self.pos = np.zeros((self.N, A, 2), np.float32)
        self.vel = np.zeros((self.N, A, 2), np.float32)
        self.lm = np.zeros((self.N, L, 2), np.float32)
        self.step_count = np.zeros(self.N, np.int32)
        self.episode = np.zeros(self.N, np.uint64)
        self._avail = np.ones((self.N, A, 5), np.float32) if self.discrete else None
        self._no_info = [[{} for _ in range(A)] for _ in range(self.N)]
        self._bad_info = [[{"bad_transition": True} for _ in range(A)] for _ in range(self.N)]

    def _reset_rows(self, rows):
        p, l = initial_positions(self.seed_value, self.env_ids[rows], self.episode[rows], self.n_agents, self.n_landmarks)
        self.pos[rows], self.lm[rows], self.vel[rows] = p, l, 0.0

    def _observe(self):
        A, L, N = self.n_agents, self.n_landmarks, self.N
        obs = np.zeros((N, A, obs_dim(A, L)), np.float32)
        for i in range(A):
            c = 0
            obs[:, i, 0:2], obs[:, i, 2:4] = self.vel[:, i], self.pos[:, i]
            c = 4
            obs[:, i, c:c + 2 * L] = (self.lm - self.pos[:, i:i + 1]).reshape(N, 2 * L)
            c += 2 * L
            others = [j for j in range(A) if j != i]
            obs[:, i, c:c + 2 * (A - 1)] = (self.pos[:, others] - self.pos[:, i:i + 1]).reshape(N, 2 * (A - 1))
        state = obs.reshape(N, 1, -1).repeat(A, axis=1)
        return obs, state

    def reset(self):
        self.episode[:] = 0
        self.step_count[:] = 0
        self._reset_rows(np.arange(self.N))
        obs, state = self._observe()
        return obs, state, self._avail

    def step(self, actions):
        A, N = self.n_agents, self.N
        actions = np.asarray(actions)
        u = np.zeros((N, A, 2), np.float64)
        if self.discrete:
            k = actions.reshape(N, A).astype(np.int64)
            u[..., 0] = np.where(k == 1, -1.0, np.where(k == 2, 1.0, 0.0))
            u[..., 1] = np.where(k == 3, -1.0, np.where(k == 4, 1.0, 0.0))
        else:
            a = actions.reshape(N, A, 5).astype(np.float64)
            u[..., 0], u[..., 1] = a[..., 1] - a[..., 2], a[..., 3] - a[..., 4]
        f = 5.0 * u
        pos = self.pos.astype(np.float64)
        for i in range(A):
            for j in range(i + 1, A):
                d = pos[:, i] - pos[:, j]
                dist = np.sqrt((d * d).sum(-1))
                pen = np.logaddexp(0.0, -(dist - 0.3) / 1e-3) * 1e-3
                fc = (1e2 * pen / dist)[:, None] * d
                f[:, i] += fc
                f[:, j] -= fc
        vel = self.vel.astype(np.float64) * 0.75 + f * 0.1
        self.vel = vel.astype(np.float32)
        self.pos = (pos + vel * 0.1).astype(np.float32)
        pos = self.pos.astype(np.float64)
        lm = self.lm.astype(np.float64)
        dist_al = np.sqrt(((pos[:, :, None, :] - lm[:, None, :, :]) ** 2).sum(-1))   # [N, A, L]
        glob = -dist_al.min(axis=1).sum(axis=1)
        dist_aa = np.sqrt(((pos[:, :, None, :] - pos[:, None, :, :]) ** 2).sum(-1))
        coll = (dist_aa < 0.3) & ~np.eye(A, dtype=bool)[None]
        local = -coll.sum(axis=2).astype(np.float64)
        team = (0.5 * glob[:, None] + 0.5 * local).sum(axis=1).astype(np.float32)
        rewards = np.repeat(team[:, None, None], A, axis=1)
        self.step_count += 1
        done = self.step_count >= self.max_cycles
        infos = self._no_info
        if done.any():
            rows = np.nonzero(done)[0]
            self.episode[rows] += np.uint64(1)
            self.step_count[rows] = 0
            self._reset_rows(rows)
            infos = self._bad_info if done.all() else [self._bad_info[n] if done[n] else self._no_info[n] for n in range(N)]
        dones = np.repeat(done[:, None], A, axis=1)
        obs, state = self._observe()
        return obs, state, rewards, dones, infos, self._avail

    def seed(self, seed):
        pass

    def close(self):
        pass


class BatchedSimpleSpread:
    """Device-resident env: one kernel per step for all rollout threads (``hb_mpe_spread_step``)."""

    team_reward = True

    def __init__(self, seed, n_threads, env_args=None, device=None, env_offset=0):
        import torch

        from .. import _lib as L_

        self._L, self._torch = L_, torch
        env_args = env_args or {}
        if env_args.get("scenario", "simple_spread_v2") != "simple_spread_v2":
            raise NotImplementedError("the native MPE backend implements simple_spread_v2 only")
        self.n_agents = A = int(env_args.get("n_agents", 3))
        self.n_landmarks = Lm = int(env_args.get("n_landmarks", A))
        if A > L_.HB_MPE_MAX_AGENTS or Lm > L_.HB_MPE_MAX_AGENTS:
            raise NotImplementedError(f"at most {L_.HB_MPE_MAX_AGENTS} agents / landmarks")
        self.n_threads = N = int(n_threads)
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("BatchedSimpleSpread needs a CUDA device (SimpleSpreadNumpy is the host twin)")
        self.discrete = not bool(env_args.get("continuous_actions", False))
        self.max_cycles = int(env_args.get("max_cycles", MAX_CYCLES))
        self.seed_value = int(seed) & (2**64 - 1)
        self.state_type = "EP"
        od = self.od = obs_dim(A, Lm)
        self.observation_space = [Box(shape=(od,)) for _ in range(A)]
        self.share_observation_space = [Box(shape=(A * od,)) for _ in range(A)]
        self.action_space = [Discrete(5) if self.discrete else Box(shape=(5,)) for _ in range(A)]
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=self.device)
        self.pos, self.vel, self.lm = z(N, A, 2), z(N, A, 2), z(N, Lm, 2)
        self.step_count = z(N, dt=torch.int32)
        self.episode = z(N, dt=torch.int64)          # uint64 on the device
        self._env_offset = int(env_offset)
        assert env_offset == 0, "sharded runs give every rank its own seed instead of an offset"
        self._obs = [z(N, od) for _ in range(A)]
        self._state, self._rew, self._dones, self._bad = z(N, A * od), z(N), z(N, A, dt=torch.uint8), z(N, A, dt=torch.uint8)
        self._avail = torch.ones(N, A, 5, device=self.device) if self.discrete else None
        self.steps_served = 0
        self.last_bad_transition = self._bad

    def _args(self, reset, actions, obs_out, share, rew, rew_na, dones, bad):
        L_ = self._L
        a = L_.MpeArgs()
        a.n_envs, a.n_agents, a.n_landmarks = self.n_threads, self.n_agents, self.n_landmarks
        a.continuous, a.max_cycles, a.reset_all, a.seed = int(not self.discrete), self.max_cycles, int(reset), self.seed_value
        a.pos, a.vel, a.landmarks = L_.ptr(self.pos), L_.ptr(self.vel), L_.ptr(self.lm)
        a.step_count, a.episode = L_.ptr(self.step_count), L_.ptr(self.episode)
        for i in range(self.n_agents):
            a.actions[i] = L_.ptr(actions[i]) if actions is not None else None
            a.obs_out[i] = L_.ptr(obs_out[i])
        a.share_obs_out, a.rewards_out, a.rewards_na_out = L_.ptr(share), L_.ptr(rew), L_.ptr(rew_na)
        a.dones_out, a.bad_out = L_.ptr(dones), L_.ptr(bad)
        return a

    def reset(self):
        a = self._args(True, None, self._obs, self._state, None, None, None, None)
        self._L.call("hb_mpe_spread_step", C.byref(a), self._L.stream_ptr())
        torch = self._torch
        obs = torch.stack(self._obs, dim=1)
        return obs, self._state.unsqueeze(1).expand(-1, self.n_agents, -1), self._avail

    def step(self, actions):
        """actions [N, A, ad] device tensor.  Returns the reference 6-tuple (device tensors, lazy infos)."""
        from .synthetic import LazyInfos

        torch = self._torch
        acts = torch.as_tensor(actions, device=self.device, dtype=torch.float32)
        per_agent = [acts[:, i].contiguous() for i in range(self.n_agents)]
        a = self._args(False, per_agent, self._obs, self._state, self._rew, None, self._dones, self._bad)
        self._L.call("hb_mpe_spread_step", C.byref(a), self._L.stream_ptr())
        self.steps_served += 1
        obs = torch.stack(self._obs, dim=1)
        share = self._state.unsqueeze(1).expand(-1, self.n_agents, -1)
        rewards = self._rew.reshape(-1, 1, 1).expand(-1, self.n_agents, -1)
        self.last_bad_transition = self._bad.bool()
        return obs, share, rewards, self._dones.bool(), LazyInfos(self._bad.bool().clone()), self._avail

    def step_into(self, dst):
        """Zero-copy step: outputs go straight into the rollout-buffer slots of ``dst`` (see SyntheticBatchedEnv)."""
        key = id(dst)
        cache = self.__dict__.setdefault("_arg_cache", {})
        a = cache.get(key)
        if a is None:
            a = cache[key] = self._args(False, dst["actions"], dst["obs"], dst["share_obs"], dst["rewards"],
                                        dst.get("rewards_na"), dst["dones"], dst["bad"])
        self._L.call("hb_mpe_spread_step", C.byref(a), self._L.stream_ptr())
        self.steps_served += 1

    def graph_period(self):
        return 1   # all per-step state lives on the device

    def graph_advance(self, steps):
        self.steps_served += steps

    def get_state(self):
        return {k: getattr(self, k).cpu().numpy() for k in ("pos", "vel", "lm", "step_count", "episode")}

    def seed(self, seed):
        pass

    def close(self):
        pass
