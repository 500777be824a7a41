"""restore() / save() / eval() of the Runner (reference on_policy_base_runner.py:500-591, 724-763).

tests/golden/ckpt_ref/*.pt were written by the UNMODIFIED reference's OnPolicyBaseRunner.save() (tests/golden/
make_golden.py `checkpoint`); tests/golden/checkpoint_ref.npz holds what the reference computes from those weights.
A checkpoint interchanges if (a) this Runner restores the reference's files and computes the same actions / values,
and (b) what this Runner saves has the reference's file names, state_dict keys, order, shapes and dtypes."""
import os
import tempfile

import numpy as np
import pytest
import torch

from tests import util as U
from tests.smoke_check import small_config

pytestmark = pytest.mark.gpu

CKPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_ref")


def _runner(model_dir=None, **train_over):
    from harl_b200.runners import RUNNER_REGISTRY

    args, algo_args, env_args = small_config()
    algo_args["train"]["model_dir"] = model_dir
    algo_args["train"].update(train_over)
    return RUNNER_REGISTRY["happo"](args, algo_args, env_args)


def test_restore_reference_checkpoint_reproduces_reference_actions_and_values():
    g = U.load("checkpoint_ref")
    r = _runner(CKPT)
    dev = r.device
    obs, avail = torch.from_numpy(g["obs"]).to(dev), torch.from_numpy(g["avail"]).to(dev)
    for a in range(3):
        act, logp, _ = r.actor[a].get_actions(obs[:, a].contiguous(), None, None, avail[:, a].contiguous(), deterministic=True)
        assert np.array_equal(act.cpu().numpy(), g[f"det_action{a}"])
        np.testing.assert_allclose(logp.cpu().numpy(), g[f"det_logp{a}"], rtol=2e-5, atol=2e-6)
    v, _ = r.critic.get_values(torch.from_numpy(g["share_obs"]).to(dev), None, None)
    np.testing.assert_allclose(v.cpu().numpy(), g["values"], rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(r.value_normalizer.state.cpu().numpy(), g["vn"])
    np.testing.assert_allclose(r.value_normalizer.denormalize(v), g["values_denorm"], rtol=2e-5, atol=2e-6)
    r.close()


def test_saved_checkpoint_has_the_reference_layout_and_round_trips():
    r = _runner(CKPT)
    r.save()
    names = sorted(os.listdir(CKPT))
    assert sorted(f for f in os.listdir(r.save_dir) if f.endswith(".pt")) == names
    for f in names:
        ref = torch.load(os.path.join(CKPT, f), map_location="cpu")
        got = torch.load(os.path.join(r.save_dir, f), map_location="cpu")   # weights_only default: plain tensors
        assert list(got.keys()) == list(ref.keys()), f
        for k in ref:
            assert got[k].shape == ref[k].shape and got[k].dtype == ref[k].dtype and got[k].device.type == "cpu", (f, k)
            assert torch.equal(got[k], ref[k]), (f, k)     # restore -> save is lossless
    # a second runner restores what the first one saved, after training moved the weights
    r.warmup()
    r.logger.init(1)
    r.run_iteration(1, 1)
    r.save()
    trained = {f: torch.load(os.path.join(r.save_dir, f), map_location="cpu") for f in names}
    assert any(not torch.equal(trained["actor_agent0.pt"][k], torch.load(os.path.join(CKPT, "actor_agent0.pt"))[k])
               for k in trained["actor_agent0.pt"])
    r2 = _runner(str(r.save_dir))
    for a in range(3):
        for k, v in r2.actor[a].actor.state_dict().items():
            assert torch.equal(v.cpu(), trained[f"actor_agent{a}.pt"][k]), k
    for k, v in r2.critic.critic.state_dict().items():
        assert torch.equal(v.cpu(), trained["critic_agent.pt"][k]), k
    assert torch.equal(r2.value_normalizer.state.cpu(), r.value_normalizer.state.cpu())
    r.close()
    r2.close()


def test_restore_fails_loudly_on_a_checkpoint_of_another_shape():
    """nn.Module.load_state_dict semantics: a missing key or a wrong shape is an error, not a silent partial load."""
    d = tempfile.mkdtemp(prefix="hb_ckpt_")
    for f in os.listdir(CKPT):
        sd = torch.load(os.path.join(CKPT, f), map_location="cpu")
        if f == "actor_agent1.pt":
            sd.pop("base.mlp.fc.3.bias")
        torch.save(sd, os.path.join(d, f))
    with pytest.raises(KeyError):
        _runner(d)
    for f in os.listdir(CKPT):
        sd = torch.load(os.path.join(CKPT, f), map_location="cpu")
        if f == "critic_agent.pt":
            sd["base.mlp.fc.0.weight"] = torch.zeros(32, 11)
        torch.save(sd, os.path.join(d, f))
    with pytest.raises((RuntimeError, ValueError)):
        _runner(d)


def test_eval_is_deterministic_and_logs_whole_episodes():
    """eval() (reference :500-591): deterministic actions on the eval envs until eval_episodes episodes finished;
    the logger's statistic is the mean return over the finished episodes.  Two runners restored from the same
    checkpoint evaluate identically; training between two evals does not consume the eval envs' episode count."""
    from harl_b200.runners import RUNNER_REGISTRY

    res = []
    for _ in range(2):
        args, algo_args, env_args = small_config()
        algo_args["train"]["model_dir"] = CKPT
        algo_args["eval"].update(use_eval=True, n_eval_rollout_threads=4, eval_episodes=9)
        r = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
        assert r.eval_envs is not None
        r.logger.init(1)
        r.logger.episode_init(1)
        r.eval()
        first = r.logger.last_eval_reward
        finished = sum(len(x) for x in r.logger.eval_episode_rewards)
        assert finished >= 9 and np.isfinite(first)
        r.eval()
        res.append((first, r.logger.last_eval_reward))
        log = open(os.path.join(r.log_dir, "progress.txt")).read().strip().splitlines() if os.path.exists(
            os.path.join(r.log_dir, "progress.txt")) else None
        if log is not None:
            assert len(log) == 2
        r.close()
    assert res[0] == res[1]
