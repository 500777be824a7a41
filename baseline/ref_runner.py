"""Time the UNMODIFIED reference (PKU-MARL/HARL, installed from /root/reference into baseline/_ref by
``pip install --no-deps --target``; git-ignored, travels to the GPU box with the gpurun snapshot) through its own
public API: ``RUNNER_REGISTRY[algo](args, algo_args, env_args).run()``.

Run as its own process (bench.py --impl reference spawns it) so that ``import harl`` resolves to baseline/_ref and not
to this repo's alias package of the same name:

    python baseline/ref_runner.py '<json spec>'

The only shims are the two SURVEY.md section 8(c) lists:
  1. a stub ``tensorboardX.SummaryWriter`` (the package is not installed; configs_tools.py:86 imports it),
  2. ``make_train_env`` of harl/runners/on_policy_base_runner.py replaced by a batched synthetic env (gym /
     pettingzoo / mujoco are not installed) -- the reference's own DexHandsEnv-style batched contract
     (harl/utils/envs_tools.py:51-54): reset() -> (obs, share_obs, avail), step(actions) -> (obs, share_obs, rewards,
     dones, infos, avail), NumPy, infos[n][a] dicts carrying "bad_transition".
Nothing of the reference is modified: its runner drives its own buffers, actors, critic, ValueNorm and logger.
Iteration times are taken from timestamps at the reference's own ``after_update`` calls (end of every iteration of
``OnPolicyBaseRunner.run``, on_policy_base_runner.py:171-267).

Prints ONE JSON object on the last stdout line.
"""
import json
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def _install_shims():
    sys.path.insert(0, REF)

    class SummaryWriter:  # configs_tools.init_dir / base_logger use these three
        scalars = []        # (tag, value dict, step) of every add_scalars call: the reference's own logged curves

        def __init__(self, *a, **k):
            pass

        def add_scalars(self, tag, values, step=None, *a, **k):
            if tag in ("train_episode_rewards", "eval_average_episode_rewards"):
                SummaryWriter.scalars.append((tag, {k_: float(v) for k_, v in values.items()}, int(step)))

        def add_scalar(self, *a, **k):
            pass

        def export_scalars_to_json(self, *a, **k):
            pass

        def close(self):
            pass

    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = SummaryWriter
    sys.modules["tensorboardX"] = tbx


class Box:  # the reference dispatches on __class__.__name__, .shape, .n only (envs_tools.py:15-46, act.py:24-34)
    def __init__(self, n):
        self.shape = (n,)

    def __eq__(self, o):
        return self.__class__ is o.__class__ and self.shape == o.shape


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()

    def __eq__(self, o):
        return self.__class__ is o.__class__ and self.n == o.n


class SyntheticEnv:
    """Host twin of harl_b200.envs.synthetic (pre-generated N(0,1) pools, truncation every ``episode_limit`` steps with
    bad_transition, optional SMAC-style deaths / terminations / availability masks), reference call contract."""

    def __init__(self, shapes, n_threads, seed=1, pool=8):
        import numpy as np

        self.np = np
        c = shapes
        self.n_agents = A = int(c["n_agents"])
        N = self.N = int(n_threads)
        od, sd, ad = int(c["obs_dim"]), int(c["share_obs_dim"]), int(c["action_dim"])
        self.discrete = c["action_type"] == "Discrete"
        self.fp = c.get("state_type", "EP") == "FP"
        self.observation_space = [Box(od) for _ in range(A)]
        self.share_observation_space = [Box(sd) for _ in range(A)]
        self.action_space = [Discrete(ad) if self.discrete else Box(ad) for _ in range(A)]
        rng = np.random.default_rng(1234 + seed)
        self.rng = rng
        self.pool = pool
        self.obs = rng.standard_normal((pool, N, A, od)).astype(np.float32)
        if self.fp:
            self.state = rng.standard_normal((pool, N, A, sd)).astype(np.float32)
        else:
            self.state = rng.standard_normal((pool, N, 1, sd)).astype(np.float32).repeat(A, axis=2)
        self.rew = rng.standard_normal((pool, N, 1, 1)).astype(np.float32).repeat(A, axis=2)
        self.death_prob = float(c.get("death_prob", 0.0))
        self.term_prob = float(c.get("terminate_prob", 0.0))
        self.avail_prob = float(c.get("avail_prob", 1.0))
        self.avail = None
        if self.discrete:
            if self.avail_prob >= 1.0:
                self.avail = np.ones((1, N, A, ad), np.float32)
            else:
                av = (rng.random((pool, N, A, ad)) < self.avail_prob).astype(np.float32)
                av[..., 0] = 0.0
                av[..., 1] = 1.0
                self.avail = av
        self.limit = int(c["episode_limit"])
        self.t = 0
        self.ep = np.zeros(N, np.int64)
        self.dead = np.zeros((N, A), bool)
        self._empty = [[{} for _ in range(A)] for _ in range(N)]

    def _avail(self, k):
        if self.avail is None:
            return None
        av = self.avail[k % self.avail.shape[0]]
        if self.death_prob > 0.0:  # dead agents: only the no-op is available (StarCraft2_Env.py:2188-2234)
            av = av.copy()
            av[self.dead] = 0.0
            av[self.dead, 0] = 1.0
        return av

    def reset(self):
        self.t = 0
        self.ep[:] = 0
        self.dead[:] = False
        return self.obs[0], self.state[0], self._avail(0)

    def step(self, actions):
        np = self.np
        self.t += 1
        self.ep += 1
        k = self.t % self.pool
        trunc = self.ep >= self.limit
        term = (self.rng.random(self.N) < self.term_prob) if self.term_prob > 0.0 else np.zeros(self.N, bool)
        done_env = trunc | term
        if self.death_prob > 0.0:
            self.dead |= self.rng.random(self.dead.shape) < self.death_prob
        dones = self.dead | done_env[:, None]
        if done_env.any():
            bad = trunc & ~term
            infos = [[({"bad_transition": True} if bad[n] else {}) for _ in range(self.n_agents)] if done_env[n] else self._empty[n]
                     for n in range(self.N)]
            self.ep[done_env] = 0
            self.dead[done_env] = False
        else:
            infos = self._empty
        return self.obs[k], self.state[k], self.rew[k], dones, infos, self._avail(k)

    def close(self):
        pass


def main():
    spec = json.loads(sys.argv[1])
    _install_shims()
    import numpy as np
    import torch

    import harl  # noqa: F401  (baseline/_ref)
    import harl.runners.on_policy_base_runner as rb
    from harl.runners import RUNNER_REGISTRY

    assert os.path.realpath(harl.__file__).startswith(os.path.realpath(REF)), harl.__file__
    shapes, N = spec["shapes"], spec["n_rollout_threads"]
    if spec.get("env_kind") == "mpe_spread":
        # the learnable task: NumPy twin of the batched simple_spread env (harl_b200/envs/mpe_spread.py); the repo root goes
        # LAST on sys.path so that `harl` keeps resolving to baseline/_ref
        sys.path.append(os.path.dirname(HERE))
        from harl_b200.envs.mpe_spread import SimpleSpreadNumpy

        rb.make_train_env = lambda env_name, seed, n_threads, env_args: SimpleSpreadNumpy(seed, n_threads, env_args)
    else:
        rb.make_train_env = lambda env_name, seed, n_threads, env_args: SyntheticEnv(shapes, n_threads, seed=1)
    args, algo_args, env_args = spec["args"], spec["algo_args"], spec["env_args"]
    W, K, budget = spec["warmup"], spec["steps"], float(spec.get("budget_s", 1e9))
    min_timed = min(K, int(spec.get("min_timed", 3)))
    T = algo_args["train"]["episode_length"]
    algo_args["train"].update(n_rollout_threads=N, num_env_steps=(W + K) * T * N, log_interval=int(spec.get("log_interval", 10**9)),
                              eval_interval=10**9)
    algo_args["eval"]["use_eval"] = False
    algo_args["device"].update(cuda=bool(spec["cuda"]), torch_threads=int(spec["torch_threads"]))
    algo_args["logger"]["log_dir"] = tempfile.mkdtemp(prefix="harl_ref_bench_")
    runner = RUNNER_REGISTRY[args["algo"]](args, algo_args, env_args)
    stamps = []
    inner = runner.after_update

    class BudgetSpent(Exception):
        pass

    def stamped():
        inner()
        if spec["cuda"]:
            torch.cuda.synchronize()
        stamps.append(time.perf_counter())
        # bounded sample: stop the reference's own loop early once the time budget is spent (>= min_timed timed iterations)
        if len(stamps) - W >= min_timed and stamps[-1] - t0 > budget:
            raise BudgetSpent()

    runner.after_update = stamped
    t0 = time.perf_counter()
    try:
        runner.run()
    except BudgetSpent:
        pass
    runner.close()
    stamps = [t0] + stamps
    per = [b - a for a, b in zip(stamps[:-1], stamps[1:])]
    timed = per[W:]
    dt = sum(timed) / len(timed)
    out = dict(value=T * N / dt, unit="env-steps/s", seconds_per_step=dt, timed_iterations=len(timed), warmup_iterations=W, iterations_s=[round(x, 4) for x in per],
               n_rollout_threads=N, episode_length=T, cuda=bool(spec["cuda"]), torch_threads=torch.get_num_threads(),
               host_cpus=os.cpu_count(), torch=torch.__version__, numpy=np.__version__,
               harl_file=os.path.relpath(harl.__file__, os.path.dirname(HERE)))
    if spec.get("log_interval"):
        out["train_episode_rewards"] = [(st, v["aver_rewards"]) for tag, v, st in sys.modules["tensorboardX"].SummaryWriter.scalars
                                        if tag == "train_episode_rewards"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
