"""Oracle: the whole on-policy iteration on the CPU (rollout -> insert -> GAE -> sequential update).

TEST INFRASTRUCTURE / CPU BASELINE (see ``oracle/__init__.py``).  This is the port
``bench.py`` times as ``cpu_baseline`` and as ``--impl reference``: it follows the reference's
control flow step for step -- host NumPy buffers, one small PyTorch forward per agent per
rollout step with host<->tensor conversions, per-iteration Python GAE loop, materialised
minibatches (including the unused rnn-state gather) -- so its cost profile is the reference's.

Restates harl/runners/on_policy_base_runner.py:171-497 (run / warmup / collect / insert /
compute / after_update) on top of oracle.algo.ha_train (on_policy_ha_runner.py:11-130) or, with
``cfg["algo_name"] == "hatrpo"``, oracle.trpo.ha_train_hatrpo; recurrent policies propagate their GRU
hidden states through the rollout (reset at episode ends) exactly as the reference runner does.
"""
import numpy as np
import torch

from . import algo as oa
from . import buffers as ob
from . import nets as on
from . import trpo as ot


class OracleRunner:
    def __init__(self, cfg, env, state_type="EP", seed=1):
        """cfg: merged model/algo/train dict with the reference's keys; env: batched env returning NumPy."""
        self.cfg, self.env, self.state_type = cfg, env, state_type
        torch.manual_seed(seed)
        self.T, self.N, self.A = cfg["episode_length"], cfg["n_rollout_threads"], env.n_agents
        T, N, A = self.T, self.N, self.A
        h, R = cfg["hidden_sizes"][-1], cfg["recurrent_n"]
        self.heads, self.actors, self.abufs = [], [], []
        for a in range(A):
            sp = env.action_space[a]
            head = sp.__class__.__name__
            od = env.observation_space[a].shape[0]
            out = sp.n if head == "Discrete" else sp.shape[0]
            ad = 1 if head == "Discrete" else out
            p = {k: v.requires_grad_(True) for k, v in on.init_params(cfg, od, head, out).items()}
            self.actors.append((p, oa.Adam(p, cfg["lr"], cfg["opti_eps"], cfg["weight_decay"])))
            self.heads.append(head)
            self.abufs.append(dict(
                obs=np.zeros((T + 1, N, od), np.float32), rnn_states=np.zeros((T + 1, N, R, h), np.float32),
                actions=np.zeros((T, N, ad), np.float32), action_log_probs=np.zeros((T, N, ad), np.float32),
                masks=np.ones((T + 1, N, 1), np.float32), active_masks=np.ones((T + 1, N, 1), np.float32),
                available_actions=np.ones((T + 1, N, out), np.float32) if head == "Discrete" else None))
        sd = env.share_observation_space[0].shape[0]
        pc = {k: v.requires_grad_(True) for k, v in on.init_params(cfg, sd, "value", 1).items()}
        self.critic = (pc, oa.Adam(pc, cfg["critic_lr"], cfg["opti_eps"], cfg["weight_decay"]))
        lead = (N,) if state_type == "EP" else (N, A)
        self.cbuf = dict(share_obs=np.zeros((T + 1, *lead, sd), np.float32),
                         rnn_states_critic=np.zeros((T + 1, *lead, R, h), np.float32),
                         value_preds=np.zeros((T + 1, *lead, 1), np.float32), returns=np.zeros((T + 1, *lead, 1), np.float32),
                         rewards=np.zeros((T, *lead, 1), np.float32), masks=np.ones((T + 1, *lead, 1), np.float32),
                         bad_masks=np.ones((T + 1, *lead, 1), np.float32))
        self.vn = ob.ValueNormState() if cfg["use_valuenorm"] else None

    def warmup(self):
        obs, share_obs, avail = self.env.reset()
        for a in range(self.A):
            self.abufs[a]["obs"][0] = obs[:, a].copy()
            if self.abufs[a]["available_actions"] is not None:
                self.abufs[a]["available_actions"][0] = avail[:, a].copy()
        self.cbuf["share_obs"][0] = share_obs[:, 0].copy() if self.state_type == "EP" else share_obs.copy()

    @torch.no_grad()
    def collect(self, step):
        """on_policy_base_runner.py:285-340."""
        acts, lps = [], []
        self._new_rnn = [None] * self.A
        for a in range(self.A):
            b, (p, _) = self.abufs[a], self.actors[a]
            obs = torch.from_numpy(b["obs"][step])
            feat, hx = on.features(p, self.cfg, obs, torch.from_numpy(b["rnn_states"][step]), torch.from_numpy(b["masks"][step]))
            self._new_rnn[a] = hx.numpy()
            if self.heads[a] == "Discrete":
                logits = on.categorical_logits(p, feat, torch.from_numpy(b["available_actions"][step]))
                act = torch.multinomial(logits.exp(), 1)
                lp = logits.gather(-1, act)
                act = act.float()
            else:
                mean, std = on.gaussian_params(p, self.cfg, feat)
                act = torch.normal(mean, std)
                lp = -((act - mean) ** 2) / (2 * std * std) - std.log() - 0.5 * on.LOG_2PI
            acts.append(act.numpy())
            lps.append(lp.numpy())
        actions = np.array(acts).transpose(1, 0, 2)
        logps = np.array(lps).transpose(1, 0, 2)
        so, rc, mk = self.cbuf["share_obs"][step], self.cbuf["rnn_states_critic"][step], self.cbuf["masks"][step]
        v, hc = on.critic_values(self.critic[0], self.cfg, torch.from_numpy(so.reshape(-1, so.shape[-1])),
                                 torch.from_numpy(rc.reshape(-1, *rc.shape[-2:])), torch.from_numpy(mk.reshape(-1, 1)))
        self._new_rnn_critic = hc.numpy().reshape(rc.shape)
        values = v.numpy().reshape(self.cbuf["value_preds"][step].shape)
        return values, actions, logps

    def insert(self, step, obs, share_obs, rewards, dones, bad, avail, values, actions, logps):
        """on_policy_base_runner.py:342-460."""
        masks, active, bad_masks, dones_env = ob.derive_masks(dones, bad, self.state_type)
        keep = (~np.asarray(dones_env, bool)).astype(np.float32)  # finished envs restart from a zero state (:358-386)
        for a in range(self.A):
            b = self.abufs[a]
            b["obs"][step + 1] = obs[:, a].copy()
            b["rnn_states"][step + 1] = self._new_rnn[a] * keep[:, None, None]
            b["actions"][step] = actions[:, a].copy()
            b["action_log_probs"][step] = logps[:, a].copy()
            b["masks"][step + 1] = masks[:, a].copy()
            b["active_masks"][step + 1] = active[:, a].copy()
            if b["available_actions"] is not None:
                b["available_actions"][step + 1] = avail[:, a].copy()
        c = self.cbuf
        ep = self.state_type == "EP"
        c["share_obs"][step + 1] = share_obs[:, 0].copy() if ep else share_obs.copy()
        c["rnn_states_critic"][step + 1] = self._new_rnn_critic * keep.reshape(-1, *([1] * (self._new_rnn_critic.ndim - 1)))
        c["value_preds"][step] = values.copy()
        c["rewards"][step] = rewards[:, 0].copy() if ep else rewards.copy()
        c["masks"][step + 1] = masks[:, 0].copy() if ep else masks.copy()
        c["bad_masks"][step + 1] = bad_masks.copy()

    @torch.no_grad()
    def compute(self):
        """on_policy_base_runner.py:462-484 + compute_returns."""
        c = self.cbuf
        so, rc, mk = c["share_obs"][-1], c["rnn_states_critic"][-1], c["masks"][-1]
        nv, _ = on.critic_values(self.critic[0], self.cfg, torch.from_numpy(so.reshape(-1, so.shape[-1])),
                                 torch.from_numpy(rc.reshape(-1, *rc.shape[-2:])), torch.from_numpy(mk.reshape(-1, 1)))
        nv = nv.numpy().reshape(c["value_preds"][-1].shape)
        ret, vp = ob.compute_returns(c["rewards"], c["value_preds"], c["masks"], c["bad_masks"], nv, self.cfg["gamma"],
                                     self.cfg["gae_lambda"], self.cfg["use_gae"], self.cfg["use_proper_time_limits"], self.vn)
        c["returns"], c["value_preds"] = ret, vp

    def after_update(self):
        for b in self.abufs:
            for k in ("obs", "rnn_states", "masks", "active_masks", "available_actions"):
                if b[k] is not None:
                    b[k][0] = b[k][-1].copy()
        for k in ("share_obs", "rnn_states_critic", "masks", "bad_masks"):
            self.cbuf[k][0] = self.cbuf[k][-1].copy()

    def run_iteration(self):
        """One reference iteration; returns (actor infos, critic info)."""
        for step in range(self.T):
            values, actions, logps = self.collect(step)
            obs, share_obs, rewards, dones, bad, avail = self.env.step(actions)
            self.insert(step, obs, share_obs, rewards, dones, bad, avail, values, actions, logps)
        self.compute()
        order = list(range(self.A)) if self.cfg["fixed_order"] else list(torch.randperm(self.A).numpy())
        perm = lambda n: torch.randperm(n).numpy()
        if self.cfg.get("algo_name") == "hatrpo":
            infos, cinfo, _, _ = ot.ha_train_hatrpo([p for p, _ in self.actors], self.critic, self.cfg, self.heads,
                                                    self.abufs, self.cbuf, self.vn, self.state_type, order, perm)
        else:
            infos, cinfo, _, _ = oa.ha_train(self.actors, self.critic, self.cfg, self.heads, self.abufs, self.cbuf, self.vn,
                                             self.state_type, order, perm)
        self.after_update()
        return infos, cinfo


class NumpySyntheticEnv:
    """Host twin of harl_b200.envs.synthetic (same shapes, same termination schedule, NumPy outputs)."""

    def __init__(self, shapes, n_threads, seed=0, pool=8):
        c = shapes
        self.n_agents = A = c["n_agents"]
        N = self.N = n_threads
        od, sd, ad = c["obs_dim"], c["share_obs_dim"], c["action_dim"]
        self.discrete = c["action_type"] == "Discrete"
        mk = lambda name, **kw: type(name, (), kw)()
        self.observation_space = [mk("Box", shape=(od,)) for _ in range(A)]
        self.share_observation_space = [mk("Box", shape=(sd,)) for _ in range(A)]
        self.action_space = [mk("Discrete", n=ad, shape=()) if self.discrete else mk("Box", shape=(ad,)) for _ in range(A)]
        rng = np.random.default_rng(1234 + seed)
        self.pool = pool
        self.obs = rng.standard_normal((pool, N, A, od)).astype(np.float32)
        self.state = rng.standard_normal((pool, N, 1, sd)).astype(np.float32).repeat(A, axis=2)
        self.rew = rng.standard_normal((pool, N, 1, 1)).astype(np.float32).repeat(A, axis=2)
        self.avail = np.ones((N, A, ad), np.float32) if self.discrete else None
        self.limit = c["episode_limit"]
        self.t = 0
        self.ep = 0

    def reset(self):
        self.t = self.ep = 0
        return self.obs[0], self.state[0], self.avail

    def step(self, actions):
        self.t += 1
        self.ep += 1
        k = self.t % self.pool
        done = self.ep >= self.limit
        if done:
            self.ep = 0
        dones = np.full((self.N, self.n_agents), done)
        return self.obs[k], self.state[k], self.rew[k], dones, dones.copy(), self.avail
