"""Learning curves of HAPPO on MPE simple_spread (3 agents, Discrete(5)): this repo on the B200 next to the UNMODIFIED
reference on the CPU, same task, same hyper-parameters (tuned_configs/pettingzoo_mpe/simple_spread_v2-discrete/happo),
same seed.

    python examples/returns_mpe.py --impl ours      --steps 1000000 --out profiles/returns_mpe_ours.json       (GPU)
    python examples/returns_mpe.py --impl reference --steps 1000000 --out profiles/returns_mpe_reference.json  (CPU)

"ours": OnPolicyHARunner of harl_b200 on the batched CUDA env (harl_b200/envs/mpe_spread.py:BatchedSimpleSpread).
"reference": baseline/_ref's OnPolicyHARunner.run() on the NumPy twin of the same world (SimpleSpreadNumpy), driven by
baseline/ref_runner.py.  Both report the reference logger's own statistic -- the mean return of the training episodes that
finished during a log interval (harl/common/base_logger.py:70-95) -- every ``--log-interval`` iterations.  Sampled actions
cannot be RNG-matched across the two (CPU torch generator vs in-kernel Philox), so the curves agree statistically, not
step for step; the worlds themselves are identical (same Philox stream for the start positions).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def config(n_threads, steps, seed):
    from harl_b200.utils.configs_tools import get_defaults_yaml_args

    algo_args, env_args = get_defaults_yaml_args("happo", "pettingzoo_mpe")
    env_args.update(scenario="simple_spread_v2", continuous_actions=False)
    algo_args["train"].update(n_rollout_threads=n_threads, episode_length=200, num_env_steps=steps, eval_interval=10**9)
    algo_args["eval"]["use_eval"] = False
    algo_args["seed"].update(seed=seed, seed_specify=True)
    algo_args["logger"]["log_dir"] = tempfile.mkdtemp(prefix="harl_returns_")
    return dict(algo="happo", env="pettingzoo_mpe", exp_name="returns"), algo_args, env_args


def run_ours(a):
    import torch

    from harl_b200.runners import RUNNER_REGISTRY

    args, algo_args, env_args = config(a.n, a.steps, a.seed)
    env_args["backend"] = "native"
    algo_args["train"]["log_interval"] = a.log_interval
    r = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
    r.disable_fast_rollout = bool(a.generic_rollout)
    r.warmup()
    T = algo_args["train"]["episode_length"]
    episodes = a.steps // T // a.n
    r.logger.init(episodes)
    curve = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for ep in range(1, episodes + 1):
        r.run_iteration(ep, episodes)
        if ep % a.log_interval == 0 and getattr(r.logger, "last_average_episode_reward", None) is not None:
            curve.append((ep * T * a.n, r.logger.last_average_episode_reward))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    r.close()
    return dict(impl="ours", device=torch.cuda.get_device_name(0), train_episode_rewards=curve, seconds=dt,
                env_steps_per_s=episodes * T * a.n / dt)


def run_reference(a):
    from harl_b200.envs.mpe_spread import obs_dim

    args, algo_args, env_args = config(a.n, a.steps, a.seed)
    T = algo_args["train"]["episode_length"]
    episodes = a.steps // T // a.n
    spec = dict(args=args, algo_args=algo_args, env_args=env_args, shapes=dict(obs_dim=obs_dim(3, 3)), env_kind="mpe_spread",
                n_rollout_threads=a.n, steps=episodes, warmup=0, cuda=False, torch_threads=a.threads, log_interval=a.log_interval,
                budget_s=1e9)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_runner.py"), json.dumps(spec)],
                       cwd=os.path.join(ROOT, "baseline"), env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr[-3000:])
    out = json.loads(r.stdout.strip().splitlines()[-1])
    return dict(impl="reference", harl_file=out["harl_file"], train_episode_rewards=out["train_episode_rewards"],
                seconds=sum(out["iterations_s"]), env_steps_per_s=out["value"], torch_threads=out["torch_threads"],
                host_cpus=out["host_cpus"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ours", "reference"], required=True)
    ap.add_argument("--steps", type=int, default=1_000_000)
    ap.add_argument("--n", type=int, default=20, help="n_rollout_threads (tuned config: 20)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--log-interval", type=int, default=5)
    ap.add_argument("--threads", type=int, default=4, help="reference: device.torch_threads (tuned config: 4)")
    ap.add_argument("--generic-rollout", action="store_true",
                    help="ours: the reference-shaped collect / step / insert loop instead of the zero-copy CUDA-graph rollout")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    res = run_ours(a) if a.impl == "ours" else run_reference(a)
    res.update(task="HAPPO pettingzoo_mpe simple_spread_v2 3 agents Discrete(5)", n_rollout_threads=a.n, episode_length=200,
               num_env_steps=a.steps, seed=a.seed, log_interval=a.log_interval,
               statistic="mean return of the training episodes finished in the log interval (base_logger.py:70-95)")
    c = res["train_episode_rewards"]
    if c:
        k = max(1, len(c) // 10)
        res["first_10pct_mean"] = sum(v for _, v in c[:k]) / k
        res["last_10pct_mean"] = sum(v for _, v in c[-k:]) / k
    txt = json.dumps(res)
    if a.out:
        with open(a.out, "w") as fh:
            fh.write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()
