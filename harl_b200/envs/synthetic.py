"""Seeded synthetic batched environment: serves pre-generated tensors of a task's shapes.

This is the measurement env of SURVEY.md section 8(d): observations / states ~ N(0,1), one
team reward broadcast to all agents, episode ends by truncation at ``episode_limit`` (with
``bad_transition``), optional random terminations, persistent agent deaths and SMAC-style
available-action masks.  It follows the reference's batched-env call contract
(harl/envs/dexhands/dexhands_env.py:29-43 is the reference's own batched env; contract in
SURVEY.md section 1 "L4 -> L0") but returns tensors resident on the training device, so the rollout
never crosses the host.  Per-step cost is a pool lookup; the pool is generated once.
"""
import torch

from .spaces import Box, Discrete


def _copy_segments(dsts, srcs, src_pinned=False, dst_pinned=False):
    """All step outputs into their buffer slots in one launch (hb_copy_segments; sources on the device or in pinned host
    memory); False -> the caller copies itself."""
    if not (dsts[0].is_cuda or srcs[0].is_cuda):
        return False
    from .. import _lib as L

    return L.copy_segments(dsts, srcs, src_pinned, dst_pinned)

# shapes of the BASELINE.json configs
PRESETS = {
    ("pettingzoo_mpe", "simple_spread_v2"): dict(n_agents=3, obs_dim=18, share_obs_dim=54, action_dim=5, episode_limit=25),
    ("mamujoco", "HalfCheetah-v2", "6x1"): dict(n_agents=6, obs_dim=23, share_obs_dim=17, action_type="Box", action_dim=1, episode_limit=1000),
    ("mamujoco", "HalfCheetah-v2", "2x3"): dict(n_agents=2, obs_dim=19, share_obs_dim=17, action_type="Box", action_dim=3, episode_limit=1000),
    ("mamujoco", "Humanoid-v2", "17x1"): dict(n_agents=17, obs_dim=393, share_obs_dim=376, action_type="Box", action_dim=1, episode_limit=1000),
    ("mamujoco", "Walker2d-v2", "6x1"): dict(n_agents=6, obs_dim=23, share_obs_dim=17, action_type="Box", action_dim=1, episode_limit=1000),
    ("smac", "5m_vs_6m"): dict(n_agents=5, obs_dim=128, share_obs_dim=128, action_type="Discrete", action_dim=12, episode_limit=70,
                               state_type="FP", death_prob=0.02, terminate_prob=0.01, avail_prob=0.6),
}


def resolve_shapes(env_name, env_args):
    """Synthetic-env parameters for a reference env name + env_args (falls back to explicit keys)."""
    cfg = dict(n_agents=3, obs_dim=18, share_obs_dim=54, action_type="Discrete", action_dim=5, state_type="EP",
               episode_limit=25, death_prob=0.0, terminate_prob=0.0, avail_prob=1.0)
    if env_name == "pettingzoo_mpe":
        cfg.update(PRESETS.get((env_name, env_args.get("scenario")), {}))
        cfg["action_type"] = "Box" if env_args.get("continuous_actions") else "Discrete"
    elif env_name == "mamujoco":
        cfg.update(PRESETS.get((env_name, env_args.get("scenario"), env_args.get("agent_conf")), {}))
    elif env_name in ("smac", "smacv2"):
        cfg.update(PRESETS.get(("smac", env_args.get("map_name")), PRESETS[("smac", "5m_vs_6m")]))
    for k in list(cfg):
        if k in env_args:
            cfg[k] = env_args[k]
    return cfg


class LazyInfos:
    """``infos[n][a]`` dicts built on demand from the last step's bad_transition flags (host copy made once)."""

    def __init__(self, bad):
        self._bad_dev, self._bad = bad, None

    def __len__(self):
        return self._bad_dev.shape[0]

    def __getitem__(self, n):
        if self._bad is None:
            self._bad = self._bad_dev.cpu().numpy()
        return [({"bad_transition": True} if b else {}) for b in self._bad[n]]


class SyntheticBatchedEnv:
    def __init__(self, env_name, seed, n_threads, env_args, device=None, pool=None):
        c = resolve_shapes(env_name, env_args)
        self.cfg = c
        # host=True: behave like a CPU simulator -- outputs live in pinned host memory and the actions
        # are pulled to the host every step (bench.py's end-to-end measurement).
        self.host = bool(env_args.get("host", False))
        self.device = torch.device("cpu" if self.host or device is None else device)
        self.n_threads = N = int(n_threads)
        self.n_agents = A = int(c["n_agents"])
        self.state_type = c["state_type"]
        od, sd, ad = int(c["obs_dim"]), int(c["share_obs_dim"]), int(c["action_dim"])
        self.discrete = c["action_type"] == "Discrete"
        self.observation_space = [Box(shape=(od,)) for _ in range(A)]
        self.share_observation_space = [Box(shape=(sd,)) for _ in range(A)]
        self.action_space = [Discrete(ad) if self.discrete else Box(shape=(ad,)) for _ in range(A)]
        self.episode_limit = int(c["episode_limit"])
        self.death_prob, self.terminate_prob, self.avail_prob = (float(c[k]) for k in ("death_prob", "terminate_prob", "avail_prob"))
        self.pool = K = int(pool if pool is not None else env_args.get("pool", 8))
        g = torch.Generator(device="cpu").manual_seed(int(seed) * 7919 + 1234)
        pin = (lambda t: t.pin_memory()) if self.host and torch.cuda.is_available() else (lambda t: t)
        rn = lambda *s: pin(torch.randn(*s, generator=g)).to(self.device)
        # agent-major pools so that obs[:, a] is contiguous for the per-agent buffers
        self._obs = rn(K, A, N, od)
        self._state = rn(K, N, sd) if self.state_type == "EP" else rn(K, N, A, sd)
        self._rew = rn(K, N, 1, 1)
        self._rand = torch.rand(K, N, A + 1, generator=g).to(self.device)
        self._avail = None
        if self.discrete:
            if self.avail_prob >= 1.0:
                self._avail = pin(torch.ones(1, N, A, ad)).to(self.device)
            else:
                av = (torch.rand(K, N, A, ad, generator=g) < self.avail_prob).float()
                av[..., 0] = 0.0  # SMAC rule: no-op unavailable while alive, "stop" always available
                av[..., 1] = 1.0
                self._avail = pin(av).to(self.device)
        self._t = 0
        self._ep_step = torch.zeros(N, dtype=torch.int32, device=self.device)
        self._dead = torch.zeros(N, A, dtype=torch.bool, device=self.device)
        self._simple = self.death_prob == 0.0 and self.terminate_prob == 0.0
        self._ep_step_host = 0
        self._true = torch.ones(N, A, dtype=torch.bool, device=self.device)
        self._false = torch.zeros(N, A, dtype=torch.bool, device=self.device)
        self.last_bad_transition = self._false
        self.steps_served = 0
        # host mode + step_into: outputs are staged in pinned host memory and copied H2D asynchronously
        self.host_staged = self.host and torch.cuda.is_available()
        self.h2d_bytes = self.d2h_bytes = 0
        self._act_host = None
        if self.host_staged:
            self._dones_h = torch.zeros(N, A, dtype=torch.uint8).pin_memory()
            self._bad_h = torch.zeros(N, A, dtype=torch.uint8).pin_memory()
            self._rew_na_h = torch.zeros(N, A).pin_memory()
            self._avail_h = torch.zeros(N, A, ad).pin_memory() if self.discrete else None

    def _views(self, k):
        obs = self._obs[k].permute(1, 0, 2)  # [N, A, od] view; obs[:, a] contiguous
        if self.state_type == "EP":
            share = self._state[k].unsqueeze(1).expand(-1, self.n_agents, -1)
        else:
            share = self._state[k]
        return obs, share

    def _avail_view(self, k, dead=None):
        if self._avail is None:
            return None
        av = self._avail[k % self._avail.shape[0]]
        if dead is not None and bool(self.death_prob > 0):
            av = av.clone()
            av[dead] = 0.0
            av[..., 0] = torch.where(dead, torch.ones_like(av[..., 0]), av[..., 0])  # dead agents: only no-op
        return av

    def reset(self):
        self._t = 0
        self._ep_step.zero_()
        self._ep_step_host = 0
        self._dead.zero_()
        obs, share = self._views(0)
        return obs, share, self._avail_view(0)

    def step(self, actions):
        """actions [N, A, ad] (ignored by the synthetic dynamics). Returns the reference 6-tuple."""
        self._t += 1
        self.steps_served += 1
        if self.host and torch.is_tensor(actions):
            self.last_actions = actions.to("cpu")  # the simulator consumes the actions on the host
        k = self._t % self.pool
        obs, share = self._views(k)
        rewards = self._rew[k].expand(-1, self.n_agents, -1)
        if self._simple:
            self._ep_step_host += 1
            if self._ep_step_host >= self.episode_limit:
                self._ep_step_host = 0
                dones, bad = self._true, self._true
            else:
                dones, bad = self._false, self._false
            avail = self._avail_view(k)
        else:
            r = self._rand[k]
            self._ep_step += 1
            trunc = self._ep_step >= self.episode_limit
            self._dead |= r[:, : self.n_agents] < self.death_prob
            env_done = trunc | self._dead.all(1) | (r[:, self.n_agents] < self.terminate_prob)
            dones = self._dead | env_done[:, None]
            bad = (trunc & env_done)[:, None].expand(-1, self.n_agents)
            self._ep_step = torch.where(env_done, torch.zeros_like(self._ep_step), self._ep_step)
            self._dead = self._dead & ~env_done[:, None]
            avail = self._avail_view(k, dones & ~env_done[:, None])
        self.last_bad_transition = bad
        return obs, share, rewards, dones, LazyInfos(bad), avail

    team_reward = True  # one reward per env broadcast to all agents (pettingzoo_mpe_env.py:56-57, StarCraft2_Env.py:681)

    def step_into(self, dst):
        """Zero-copy variant of ``step``: write this step's outputs straight into the rollout-buffer slots.

        ``dst``: obs (list per agent, [N, od]), share_obs, rewards (critic slot), avail (list or None entries),
        actions (list per agent -- ignored by the synthetic dynamics), dones / bad ([N, A] uint8)."""
        if self.host:
            return self._step_into_from_host(dst)
        self._t += 1
        self.steps_served += 1
        k = self._t % self.pool
        A = self.n_agents
        # one multi-tensor copy (a single launch) instead of A + 2 separate copy kernels
        dsts = list(dst["obs"]) + [dst["share_obs"]]
        srcs = [self._obs[k, a] for a in range(A)] + [self._state[k]]
        rew = self._rew[k, :, 0]
        if self.state_type == "EP":
            dsts.append(dst["rewards"])
            srcs.append(rew)
        else:
            dst["rewards"].copy_(rew.unsqueeze(1).expand(-1, A, -1))
        if dst.get("rewards_na") is not None:
            dst["rewards_na"].copy_(rew.expand(-1, A))
        if not _copy_segments(dsts, srcs):
            torch._foreach_copy_(dsts, srcs)
        if self._simple:
            self._ep_step_host += 1
            done = self._ep_step_host >= self.episode_limit
            if done:
                self._ep_step_host = 0
            if getattr(self, "_last_done_written", None) is not done or dst["dones"].data_ptr() != getattr(self, "_last_done_ptr", 0):
                dst["dones"].fill_(1 if done else 0)
                dst["bad"].fill_(1 if done else 0)
                self._last_done_written, self._last_done_ptr = done, dst["dones"].data_ptr()
            if self._avail is not None and self._avail.shape[0] > 1:
                av = self._avail_view(k)
                for a in range(A):
                    dst["avail"][a].copy_(av[:, a])
        else:
            r = self._rand[k]
            self._ep_step += 1
            trunc = self._ep_step >= self.episode_limit
            self._dead |= r[:, :A] < self.death_prob
            env_done = trunc | self._dead.all(1) | (r[:, A] < self.terminate_prob)
            dones = self._dead | env_done[:, None]
            dst["dones"].copy_(dones)
            dst["bad"].copy_((trunc & env_done)[:, None].expand(-1, A))
            self._ep_step.masked_fill_(env_done, 0)          # in place: the state must carry across CUDA-graph replays
            self._dead &= ~env_done[:, None]
            av = self._avail_view(k, dones & ~env_done[:, None])
            if av is not None:
                for a in range(A):
                    dst["avail"][a].copy_(av[:, a])
            self.last_bad_transition = dst["bad"].bool()

    def _step_into_from_host(self, dst):
        """``step_into`` for the host-resident env: the actions come D2H (the simulator consumes them before it
        steps), every output goes H2D from pinned memory straight into the buffer slots.  One stream
        synchronisation per step -- the true data dependency of a CPU simulator."""
        A = self.n_agents
        if self._act_host is None:
            self._act_host = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in dst["actions"]]
        # all agents' actions in one launch (the kernel writes the pinned block over PCIe) instead of A D2H DMA set-ups
        if not (getattr(self, "_act_zero_copy_ok", True) and _copy_segments(self._act_host, list(dst["actions"]), dst_pinned=True)):
            self._act_zero_copy_ok = False
            for a in range(A):
                self._act_host[a].copy_(dst["actions"][a], non_blocking=True)
        self.d2h_bytes += sum(x.numel() * 4 for x in self._act_host)
        torch.cuda.current_stream().synchronize()  # also: last step's H2D copies have drained the staging buffers
        self.last_actions = self._act_host
        self._t += 1
        self.steps_served += 1
        k = self._t % self.pool
        if self._simple:
            return self._step_into_packed(dst, k)

        def h2d(dst_t, src):
            dst_t.copy_(src, non_blocking=True)
            self.h2d_bytes += dst_t.numel() * dst_t.element_size()

        for a in range(A):
            h2d(dst["obs"][a], self._obs[k, a])
        h2d(dst["share_obs"], self._state[k])
        rew = self._rew[k, :, 0]
        if self.state_type == "EP":
            h2d(dst["rewards"], rew)
        else:
            self._rew_na_h.copy_(rew.expand(-1, A))
            h2d(dst["rewards"], self._rew_na_h.unsqueeze(-1))
        if dst.get("rewards_na") is not None:
            self._rew_na_h.copy_(rew.expand(-1, A))
            h2d(dst["rewards_na"], self._rew_na_h)
        if self._simple:
            self._ep_step_host += 1
            done = self._ep_step_host >= self.episode_limit
            if done:
                self._ep_step_host = 0
            self._dones_h.fill_(1 if done else 0)
            self._bad_h.fill_(1 if done else 0)
            av = self._avail_view(k)
        else:
            r = self._rand[k]
            self._ep_step += 1
            trunc = self._ep_step >= self.episode_limit
            self._dead |= r[:, :A] < self.death_prob
            env_done = trunc | self._dead.all(1) | (r[:, A] < self.terminate_prob)
            dones = self._dead | env_done[:, None]
            self._dones_h.copy_(dones)
            self._bad_h.copy_((trunc & env_done)[:, None].expand(-1, A))
            self._ep_step = torch.where(env_done, torch.zeros_like(self._ep_step), self._ep_step)
            self._dead = self._dead & ~env_done[:, None]
            av = self._avail_view(k, dones & ~env_done[:, None])
            self.last_bad_transition = self._bad_h.bool()
        h2d(dst["dones"], self._dones_h)
        h2d(dst["bad"], self._bad_h)
        if av is not None:
            self._avail_h.copy_(av)
            for a in range(A):
                h2d(dst["avail"][a], self._avail_h[:, a])

    def _step_into_packed(self, dst, k):
        """Simple dynamics, host-resident: the simulator's outputs of one step sit packed in ONE pinned block
        (agent-major obs | state | reward | avail), so a step costs one large H2D copy into a device staging block,
        two small ones (dones / bad-transition flags) and one multi-tensor device copy into the buffer slots --
        instead of 3A + 4 separate DMA set-ups."""
        A, N = self.n_agents, self.n_threads
        if getattr(self, "_pack", None) is None:
            od = self._obs.shape[-1]
            st = self._state[0].numel()
            ad = self._avail.shape[-1] if self._avail is not None else 0
            P = A * N * od + st + N + A * N * ad
            pack = torch.empty(self.pool, P).pin_memory()
            o0, o1, o2 = A * N * od, A * N * od + st, A * N * od + st + N
            for kk in range(self.pool):
                pack[kk, :o0] = self._obs[kk].reshape(-1)
                pack[kk, o0:o1] = self._state[kk].reshape(-1)
                pack[kk, o1:o2] = self._rew[kk].reshape(-1)
                if ad:
                    pack[kk, o2:] = self._avail[kk % self._avail.shape[0]].permute(1, 0, 2).reshape(-1)  # agent-major
            dev = dst["share_obs"].device
            self._pack, self._pack_d = pack, torch.empty(P, device=dev)
            d = self._pack_d
            views = lambda d: ([d[a * N * od:(a + 1) * N * od].view(N, od) for a in range(A)], d[o0:o1].view(self._state[0].shape),
                               d[o1:o2].view(N, 1), [d[o2 + a * N * ad:o2 + (a + 1) * N * ad].view(N, ad) for a in range(A)] if ad else None)
            self._pack_views = views(self._pack_d)
            self._pack_host_views = [views(pack[kk]) for kk in range(self.pool)]   # the same layout, in the pinned block
        self._ep_step_host += 1
        done = self._ep_step_host >= self.episode_limit
        if done:
            self._ep_step_host = 0
        self._dones_h.fill_(1 if done else 0)
        self._bad_h.fill_(1 if done else 0)
        # Zero-copy path: ONE kernel (hb_copy_segments) reads the pinned host block over PCIe and writes every output --
        # observations, state, rewards, availability masks, done and bad-transition flags -- straight into its buffer slot.
        # (Pinned memory is device-accessible under unified addressing.)  One launch instead of three H2D DMA set-ups plus a
        # device-side scatter; the bytes that cross PCIe are the same.
        if self.state_type == "EP" and dst.get("rewards_na") is None and getattr(self, "_zero_copy_ok", True):
            hv = self._pack_host_views[k]
            dsts = list(dst["obs"]) + [dst["share_obs"], dst["rewards"]] + (list(dst["avail"]) if hv[3] is not None else []) + \
                [dst["dones"], dst["bad"]]
            srcs = list(hv[0]) + [hv[1], hv[2]] + (list(hv[3]) if hv[3] is not None else []) + [self._dones_h, self._bad_h]
            if _copy_segments(dsts, srcs, src_pinned=True):
                self.h2d_bytes += self._pack_d.numel() * 4 + 2 * self._dones_h.numel()
                return
            self._zero_copy_ok = False   # misaligned shapes: the staged path below
        self._pack_d.copy_(self._pack[k], non_blocking=True)
        self.h2d_bytes += self._pack_d.numel() * 4
        dst["dones"].copy_(self._dones_h, non_blocking=True)
        dst["bad"].copy_(self._bad_h, non_blocking=True)
        self.h2d_bytes += 2 * self._dones_h.numel()
        obs_v, state_v, rew_v, avail_v = self._pack_views
        dsts, srcs = list(dst["obs"]) + [dst["share_obs"]], list(obs_v) + [state_v]
        if self.state_type == "EP":
            dsts.append(dst["rewards"])
            srcs.append(rew_v)
        else:
            dst["rewards"].copy_(rew_v.unsqueeze(1).expand(-1, A, -1))
        if dst.get("rewards_na") is not None:
            dst["rewards_na"].copy_(rew_v.expand(-1, A))
        if avail_v is not None:
            dsts += list(dst["avail"])
            srcs += list(avail_v)
        if not _copy_segments(dsts, srcs):
            torch._foreach_copy_(dsts, srcs)

    def graph_period(self):
        """Number of steps after which the HOST side of ``step_into`` repeats itself (pool index, and in the simple
        mode the host episode counter) -- a rollout of a multiple of this many steps can be captured into a CUDA
        graph and replayed.  None: not capturable (host-resident env)."""
        if self.host:
            return None
        if self._simple:
            import math

            return self.pool * self.episode_limit // math.gcd(self.pool, self.episode_limit)
        return self.pool

    def graph_advance(self, steps):
        """Bookkeeping for a replayed graph of ``steps`` steps (the device state advanced inside the graph)."""
        self._t += steps
        self.steps_served += steps
        if self._simple:
            self._ep_step_host = (self._ep_step_host + steps) % self.episode_limit

    def seed(self, seed):
        pass

    def close(self):
        pass
