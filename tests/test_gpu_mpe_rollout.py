"""The zero-copy rollout on the native MPE env (hb_rollout_collect + BatchedSimpleSpread.step_into + the insert kernel,
replayed from a CUDA graph) must fill the buffers exactly like the reference-shaped collect / step / insert loop on the
same env, and both must see the same world as the NumPy twin."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _runner(fast, n=16, T=60):
    import tempfile

    from harl_b200.runners import RUNNER_REGISTRY
    from harl_b200.utils.configs_tools import get_defaults_yaml_args

    algo_args, env_args = get_defaults_yaml_args("happo", "pettingzoo_mpe")
    env_args.update(scenario="simple_spread_v2", continuous_actions=False, backend="native")
    algo_args["train"].update(n_rollout_threads=n, episode_length=T, num_env_steps=10**9, log_interval=10**9, eval_interval=10**9)
    algo_args["eval"]["use_eval"] = False
    algo_args["algo"]["fixed_order"] = True
    algo_args["logger"]["log_dir"] = tempfile.mkdtemp(prefix="hb_mpe_")
    r = RUNNER_REGISTRY["happo"](dict(algo="happo", env="pettingzoo_mpe", exp_name="t"), algo_args, env_args)
    r.disable_fast_rollout = not fast
    r.warmup()
    r.logger.init(10**9)
    return r


def _snap(r):
    out = {}
    for a in range(r.num_agents):
        b = r.actor_buffer[a]
        for k in ("obs", "actions", "action_log_probs", "masks", "active_masks", "available_actions"):
            if getattr(b, k) is not None:
                out[f"a{a}.{k}"] = getattr(b, k).clone()
    for k in ("share_obs", "value_preds", "returns", "rewards", "masks", "bad_masks", "advantages"):
        out["c." + k] = getattr(r.critic_buffer, k).clone()
    return out


def test_native_env_fast_rollout_equals_generic_rollout_and_twin():
    from harl_b200 import _lib as L
    from harl_b200.envs.mpe_spread import SimpleSpreadNumpy

    f, g = _runner(True), _runner(False)
    assert f.envs.seed_value == g.envs.seed_value
    for it in (1, 2):
        for r in (f, g):
            r.run_iteration(it, 10**9)
        torch.cuda.synchronize()
        assert bool(f._fast) and not bool(g._fast)
        if it == 2:
            break   # the second iteration starts from weights that differ in the last bits
        sf, sg = _snap(f), _snap(g)
        exact = L.lib.hb_get_gemm_impl() == 0
        for k in sf:
            kind = k.split(".")[1]
            if kind in ("masks", "active_masks", "available_actions", "bad_masks"):
                assert torch.equal(sf[k], sg[k]), k
            elif exact:
                assert torch.equal(sf[k], sg[k]), k
        # the episode structure: 60 steps = 2 whole episodes of 25 + 10 steps; truncation -> masks 0, bad_masks 0
        cm, cb = sf["c.masks"][:, :, 0].cpu().numpy(), sf["c.bad_masks"][:, :, 0].cpu().numpy()
        assert (cm[[25, 50]] == 0).all() and (cb[[25, 50]] == 0).all()
        assert cm.sum() == cm.size - 2 * cm.shape[1] and cb.sum() == cb.size - 2 * cb.shape[1]
        # replay the fast runner's actions on the NumPy twin: same observations, rewards, share_obs, step by step
        tw = SimpleSpreadNumpy(f.envs.seed_value, 16, {})
        o, s, _ = tw.reset()
        T = sf["a0.actions"].shape[0]
        # contacts are stiff (force 1e2 * softplus(-(d - 0.3) / 1e-3)): a last-bit difference between libm and CUDA exp/log
        # can grow inside an episode, so a few entries may differ visibly; everything else agrees to rounding
        close = lambda x, y, tol: float((np.abs(x - y) <= tol).mean())
        for t in range(T):
            for a in range(3):
                assert close(sf[f"a{a}.obs"][t].cpu().numpy(), o[:, a], 1e-4) > 0.98, f"obs t={t} agent {a}"
            assert close(sf["c.share_obs"][t].cpu().numpy(), s[:, 0], 1e-4) > 0.98, f"share_obs t={t}"
            acts = np.stack([sf[f"a{a}.actions"][t].cpu().numpy() for a in range(3)], axis=1)
            o, s, rew, dones, infos, _ = tw.step(acts)
            assert close(sf["c.rewards"][t].cpu().numpy(), rew[:, 0], 1e-3) > 0.9, f"reward t={t}"
            assert dones.all() == ((t + 1) % 25 == 0)
    for r in (f, g):
        r.close()
