"""Per-agent rollout storage, device-resident (reference: harl/common/buffers/on_policy_actor_buffer.py).

Same attribute names, shapes ([T+1, N, ...] / [T, N, ...], time-major) and slot semantics as
the reference (SURVEY.md Appendix A "slot timeline"), but the arrays are torch tensors on the
training device for the whole run: the kernels read them in place, nothing is staged through
the host.  Non-recurrent policies never touch ``rnn_states``; it is then a zero-stride view
(the reference allocates 421 MB per agent at C2 for it and gathers it every minibatch).

The three minibatch generators are kept for API compatibility and for the index-map parity
tests (they yield the reference's NumPy tuples); the device training path does not materialise
minibatches -- it passes index tensors to the kernels.
"""
import numpy as np
import torch

from ...utils.envs_tools import get_shape_from_act_space, get_shape_from_obs_space
from ...utils.trans_tools import _flatten


def _as_tensor(x, like):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=like.device, dtype=like.dtype)


class OnPolicyActorBuffer:
    def __init__(self, args, obs_space, act_space, device=torch.device("cpu")):
        self.episode_length = T = args["episode_length"]
        self.n_rollout_threads = N = args["n_rollout_threads"]
        self.hidden_sizes = args["hidden_sizes"]
        self.rnn_hidden_size = self.hidden_sizes[-1]
        self.recurrent_n = args["recurrent_n"]
        self.device = torch.device(device)
        self.recurrent = bool(args.get("use_recurrent_policy") or args.get("use_naive_recurrent_policy"))
        obs_shape = get_shape_from_obs_space(obs_space)
        if isinstance(obs_shape[-1], list):
            obs_shape = obs_shape[:1]
        kw = dict(dtype=torch.float32, device=self.device)
        self.obs = torch.zeros(T + 1, N, *obs_shape, **kw)
        if self.recurrent:
            self.rnn_states = torch.zeros(T + 1, N, self.recurrent_n, self.rnn_hidden_size, **kw)
        else:
            self.rnn_states = torch.zeros(1, 1, self.recurrent_n, self.rnn_hidden_size, **kw).expand(
                T + 1, N, self.recurrent_n, self.rnn_hidden_size)
        if act_space.__class__.__name__ == "Discrete":
            self.available_actions = torch.ones(T + 1, N, act_space.n, **kw)
        else:
            self.available_actions = None
        act_shape = get_shape_from_act_space(act_space)
        self.actions = torch.zeros(T, N, act_shape, **kw)
        self.action_log_probs = torch.zeros(T, N, act_shape, **kw)
        self.masks = torch.ones(T + 1, N, 1, **kw)
        self.active_masks = torch.ones(T + 1, N, 1, **kw)
        self.factor = None
        self.step = 0

    def update_factor(self, factor):
        """Save the running importance-ratio product for this agent (reference :78-80)."""
        self.factor = _as_tensor(factor, self.masks).clone()

    def insert(self, obs, rnn_states, actions, action_log_probs, masks, active_masks=None, available_actions=None):
        """Write slot step+1 (obs, rnn, masks, active, avail) and slot step (actions, log-probs) (reference :82-103).

        Arguments that already alias their destination slot (the zero-copy rollout path) are skipped."""
        s = self.step

        def put(dst, src):
            if src is None:
                return
            src = _as_tensor(src, dst)
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src.reshape(dst.shape))

        put(self.obs[s + 1], obs)
        if self.recurrent:
            put(self.rnn_states[s + 1], rnn_states)
        put(self.actions[s], actions)
        put(self.action_log_probs[s], action_log_probs)
        put(self.masks[s + 1], masks)
        put(self.active_masks[s + 1], active_masks)
        if self.available_actions is not None:
            put(self.available_actions[s + 1], available_actions)
        self.step = (s + 1) % self.episode_length

    def after_update(self):
        """Slot T -> slot 0 (reference :105-112)."""
        self.obs[0].copy_(self.obs[-1])
        if self.recurrent:
            self.rnn_states[0].copy_(self.rnn_states[-1])
        self.masks[0].copy_(self.masks[-1])
        self.active_masks[0].copy_(self.active_masks[-1])
        if self.available_actions is not None:
            self.available_actions[0].copy_(self.available_actions[-1])

    # ------------------------------------------------------------------ reference-compatible generators
    def _np(self, x):
        return x.detach().cpu().numpy()

    def _pack(self, obs, rnn, act, masks, active, lp, adv, avail, factor):
        out = [self._np(obs), self._np(rnn), self._np(act), self._np(masks), self._np(active), self._np(lp),
               None if adv is None else self._np(adv), None if avail is None else self._np(avail)]
        if factor is not None:
            out.append(self._np(factor))
        return tuple(out)

    def feed_forward_generator_actor(self, advantages, actor_num_mini_batch=None, mini_batch_size=None):
        """Reference :114-178: row k of the time-major flatten <-> (t = k // N, n = k % N)."""
        T, N = self.actions.shape[:2]
        batch = T * N
        if mini_batch_size is None:
            assert batch >= actor_num_mini_batch
            mini_batch_size = batch // actor_num_mini_batch
        rand = torch.randperm(batch)
        adv = None if advantages is None else _as_tensor(advantages, self.masks).reshape(-1, 1)
        fl = lambda a: a.reshape(batch, *a.shape[2:])
        for i in range(actor_num_mini_batch):
            idx = rand[i * mini_batch_size:(i + 1) * mini_batch_size].to(self.device)
            yield self._pack(fl(self.obs[:-1])[idx], fl(self.rnn_states[:-1])[idx], fl(self.actions)[idx],
                             fl(self.masks[:-1])[idx], fl(self.active_masks[:-1])[idx], fl(self.action_log_probs)[idx],
                             None if adv is None else adv[idx],
                             None if self.available_actions is None else fl(self.available_actions[:-1])[idx],
                             None if self.factor is None else fl(self.factor)[idx])

    def naive_recurrent_generator_actor(self, advantages, actor_num_mini_batch):
        """Reference :180-221: whole trajectories of N // num_mini_batch envs per batch."""
        T, N = self.actions.shape[:2]
        assert N >= actor_num_mini_batch
        k = N // actor_num_mini_batch
        perm = torch.randperm(N)
        adv = _as_tensor(advantages, self.masks)
        for i in range(actor_num_mini_batch):
            ids = perm[i * k:(i + 1) * k].to(self.device)
            f = lambda a: _flatten(T, k, a[:T, ids])
            yield self._pack(f(self.obs), self.rnn_states[0, ids], f(self.actions), f(self.masks), f(self.active_masks),
                             f(self.action_log_probs), f(adv),
                             None if self.available_actions is None else f(self.available_actions),
                             None if self.factor is None else f(self.factor))

    def recurrent_generator_actor(self, advantages, actor_num_mini_batch, data_chunk_length):
        """Reference :223-326: env-major chunks of length L, hidden state stored at each chunk start."""
        T, N = self.actions.shape[:2]
        L = data_chunk_length
        chunks = (T * N) // L
        mb = chunks // actor_num_mini_batch
        assert T % L == 0 and chunks >= 2
        rand = torch.randperm(chunks)
        per_env = T // L
        adv = _as_tensor(advantages, self.masks)
        steps = torch.arange(L, device=self.device)
        for i in range(actor_num_mini_batch):
            c = rand[i * mb:(i + 1) * mb].to(self.device)
            n, t0 = c // per_env, (c % per_env) * L
            tt = (t0[None, :] + steps[:, None]).reshape(-1)
            nn = n.repeat(L)
            g = lambda a: a[tt, nn]
            yield self._pack(g(self.obs), self.rnn_states[t0, n], g(self.actions), g(self.masks), g(self.active_masks),
                             g(self.action_log_probs), g(adv),
                             None if self.available_actions is None else g(self.available_actions),
                             None if self.factor is None else g(self.factor))
