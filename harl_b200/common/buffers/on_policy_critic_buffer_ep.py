"""Shared-state critic rollout storage, device-resident
(reference: harl/common/buffers/on_policy_critic_buffer_ep.py; the FP twin adds an agent axis).

``compute_returns`` launches the GAE / return-scan kernel (hb_gae_returns), which also
produces the un-normalised advantages the sequential-agent update starts from
(on_policy_ha_runner.py:26-33) -- they are kept in ``self.advantages``.
"""
import numpy as np
import torch

from ... import _lib as L
from ...utils.envs_tools import get_shape_from_obs_space
from ...utils.trans_tools import _flatten


def _as_tensor(x, like):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=like.device, dtype=like.dtype)


class OnPolicyCriticBufferEP:
    """EP = "environment provided" global state, identical for all agents: arrays are [T(+1), N, ...]."""

    def __init__(self, args, share_obs_space, num_agents=None, device=torch.device("cpu")):
        self.episode_length = T = args["episode_length"]
        self.n_rollout_threads = N = args["n_rollout_threads"]
        self.hidden_sizes = args["hidden_sizes"]
        self.rnn_hidden_size = self.hidden_sizes[-1]
        self.recurrent_n = args["recurrent_n"]
        self.gamma = args["gamma"]
        self.gae_lambda = args["gae_lambda"]
        self.use_gae = args["use_gae"]
        self.use_proper_time_limits = args["use_proper_time_limits"]
        self.device = torch.device(device)
        self.num_agents = num_agents
        self.recurrent = bool(args.get("use_recurrent_policy") or args.get("use_naive_recurrent_policy"))
        lead = (N,) if num_agents is None else (N, num_agents)
        shp = get_shape_from_obs_space(share_obs_space)
        if isinstance(shp[-1], list):
            shp = shp[:1]
        kw = dict(dtype=torch.float32, device=self.device)
        self.share_obs = torch.zeros(T + 1, *lead, *shp, **kw)
        if self.recurrent:
            self.rnn_states_critic = torch.zeros(T + 1, *lead, self.recurrent_n, self.rnn_hidden_size, **kw)
        else:
            self.rnn_states_critic = torch.zeros(*([1] * (1 + len(lead))), self.recurrent_n, self.rnn_hidden_size,
                                                 **kw).expand(T + 1, *lead, self.recurrent_n, self.rnn_hidden_size)
        self.value_preds = torch.zeros(T + 1, *lead, 1, **kw)
        self.returns = torch.zeros(T + 1, *lead, 1, **kw)
        self.rewards = torch.zeros(T, *lead, 1, **kw)
        self.masks = torch.ones(T + 1, *lead, 1, **kw)
        self.bad_masks = torch.ones(T + 1, *lead, 1, **kw)
        self.advantages = torch.zeros(T, *lead, 1, **kw)
        self.step = 0

    def insert(self, share_obs, rnn_states_critic, value_preds, rewards, masks, bad_masks):
        """Reference :73-84.  Arguments aliasing their destination slot are skipped (zero-copy rollout)."""
        s = self.step

        def put(dst, src):
            if src is None:
                return
            src = _as_tensor(src, dst)
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src.reshape(dst.shape))

        put(self.share_obs[s + 1], share_obs)
        if self.recurrent:
            put(self.rnn_states_critic[s + 1], rnn_states_critic)
        put(self.value_preds[s], value_preds)
        put(self.rewards[s], rewards)
        put(self.masks[s + 1], masks)
        put(self.bad_masks[s + 1], bad_masks)
        self.step = (s + 1) % self.episode_length

    def after_update(self):
        """Slot T -> slot 0 for share_obs, rnn states, masks, bad_masks (reference :86-91)."""
        self.share_obs[0].copy_(self.share_obs[-1])
        if self.recurrent:
            self.rnn_states_critic[0].copy_(self.rnn_states_critic[-1])
        self.masks[0].copy_(self.masks[-1])
        self.bad_masks[0].copy_(self.bad_masks[-1])

    def get_mean_rewards(self):
        return float(self.rewards.mean().item())

    def compute_returns(self, next_value, value_normalizer=None):
        """Reference :97-200 -- all four use_gae x use_proper_time_limits branches, with or without ValueNorm."""
        if self.device.type != "cuda":
            raise RuntimeError("compute_returns runs the CUDA GAE kernel (no CPU fallback)")
        nv = _as_tensor(next_value, self.returns).reshape(-1).contiguous()
        T = self.episode_length
        C = self.value_preds[0].numel()
        vn = value_normalizer.state if value_normalizer is not None else None
        L.call("hb_gae_returns", L.ptr(self.rewards), L.ptr(self.value_preds), L.ptr(self.masks), L.ptr(self.bad_masks),
               L.ptr(nv), L.ptr(self.returns), L.ptr(self.advantages), T, C, float(np.float32(self.gamma)),
               float(np.float32(self.gamma * self.gae_lambda)), int(bool(self.use_gae)),
               int(bool(self.use_proper_time_limits)), L.ptr(vn), L.stream_ptr())

    # ------------------------------------------------------------------ reference-compatible generators
    def _rows(self, a):
        """[T(+1), N(, A), ...] -> [T(+1), C, ...] with C = N or N*A (agent fastest, as _ma_cast / FP flatten)."""
        return a.reshape(a.shape[0], -1, *a.shape[1 + (1 if self.num_agents is None else 2):])

    def _pack(self, *xs):
        return tuple(x.detach().cpu().numpy() for x in xs)

    def feed_forward_generator_critic(self, critic_num_mini_batch=None, mini_batch_size=None):
        """Reference :202-250 (FP :212-256): time-major flatten, k -> (t, n[, a])."""
        so, rn, vp, rt, mk = (self._rows(a) for a in (self.share_obs, self.rnn_states_critic, self.value_preds,
                                                       self.returns, self.masks))
        T, Cn = self.episode_length, vp.shape[1]
        batch = T * Cn
        if mini_batch_size is None:
            assert batch >= critic_num_mini_batch
            mini_batch_size = batch // critic_num_mini_batch
        rand = torch.randperm(batch)
        fl = lambda a: a[:T].reshape(batch, *a.shape[2:])
        for i in range(critic_num_mini_batch):
            idx = rand[i * mini_batch_size:(i + 1) * mini_batch_size].to(self.device)
            yield self._pack(fl(so)[idx], fl(rn)[idx], fl(vp)[idx], fl(rt)[idx], fl(mk)[idx])

    def naive_recurrent_generator_critic(self, critic_num_mini_batch):
        """Reference :252-283: whole trajectories; FP permutes the N*A (env, agent) pairs."""
        so, rn, vp, rt, mk = (self._rows(a) for a in (self.share_obs, self.rnn_states_critic, self.value_preds,
                                                       self.returns, self.masks))
        T, Cn = self.episode_length, vp.shape[1]
        assert Cn >= critic_num_mini_batch
        k = Cn // critic_num_mini_batch
        perm = torch.randperm(Cn)
        for i in range(critic_num_mini_batch):
            ids = perm[i * k:(i + 1) * k].to(self.device)
            f = lambda a: _flatten(T, k, a[:T, ids])
            yield self._pack(f(so), rn[0, ids], f(vp), f(rt), f(mk))

    def recurrent_generator_critic(self, critic_num_mini_batch, data_chunk_length):
        """Reference :285-369 (FP :306-390): chunks of length L per (env[, agent]) column."""
        so, rn, vp, rt, mk = (self._rows(a) for a in (self.share_obs, self.rnn_states_critic, self.value_preds,
                                                       self.returns, self.masks))
        T, Cn = self.episode_length, vp.shape[1]
        Lc = data_chunk_length
        chunks = (T * Cn) // Lc
        mb = chunks // critic_num_mini_batch
        assert T % Lc == 0 and chunks >= 2
        rand = torch.randperm(chunks)
        per = T // Lc
        steps = torch.arange(Lc, device=self.device)
        for i in range(critic_num_mini_batch):
            c = rand[i * mb:(i + 1) * mb].to(self.device)
            n, t0 = c // per, (c % per) * Lc
            tt = (t0[None, :] + steps[:, None]).reshape(-1)
            nn = n.repeat(Lc)
            g = lambda a: a[tt, nn]
            yield self._pack(g(so), rn[t0, n], g(vp), g(rt), g(mk))
