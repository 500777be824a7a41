"""The zero-copy rollout on the native MPE env (hb_rollout_collect + BatchedSimpleSpread.step_into + the insert kernel,
replayed from a CUDA graph) must fill the buffers exactly like the reference-shaped collect / step / insert loop on the
same env (the env itself is pinned to its NumPy twin in tests/test_mpe_spread.py)."""
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


def test_native_env_fast_rollout_equals_generic_rollout():
    from harl_b200 import _lib as L

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
        # the episode structure: every 25th step is a truncation -> masks 0 and bad_masks 0 in the next slot, all envs at once
        cm, cb = sf["c.masks"][1:, :, 0].cpu().numpy(), sf["c.bad_masks"][1:, :, 0].cpu().numpy()
        ends = np.where(cm[:, 0] == 0)[0]
        assert len(ends) in (2, 3) and (np.diff(ends) == 25).all()
        assert (cm[ends] == 0).all() and (cb[ends] == 0).all()
        assert cm.sum() == cm.size - len(ends) * cm.shape[1] and cb.sum() == cb.size - len(ends) * cb.shape[1]
    for r in (f, g):
        r.close()
