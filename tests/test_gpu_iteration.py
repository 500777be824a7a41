"""GPU end-to-end parity: a whole training iteration through the public Runner vs the CPU oracle,
and the reference's own ``OnPolicyHARunner.train()`` golden vectors replayed on the device.

Tolerances: masks / returns / advantages bit-exact; factors 3e-4 rel; train-info scalars 2e-4;
weights after the update 3e-5 abs (Adam steps are lr-sized, 5e-4).
"""
import numpy as np
import pytest
import torch

from tests import util as U
from tests.smoke_check import check_iteration, small_config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("action_type,state_type,over", [
    ("Discrete", "EP", {}),
    ("Box", "EP", dict(clip_param=0.05)),
    ("Discrete", "FP", dict(gamma=0.95)),
    ("Box", "FP", dict(action_aggregation="mean")),
    ("Discrete", "EP", dict(use_policy_active_masks=False, use_huber_loss=False, use_clipped_value_loss=False)),
])
def test_iteration_vs_oracle(action_type, state_type, over):
    from harl_b200.runners import RUNNER_REGISTRY

    args, algo_args, env_args = small_config(action_type=action_type, state_type=state_type, **over)
    runner = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
    runner.warmup()
    runner.logger.init(2)
    check_iteration(runner)
    check_iteration(runner)  # second iteration: slot T -> 0 carry-over, Adam moments, ValueNorm state
    runner.close()


@pytest.mark.parametrize("hidden,obs_dim,action_type", [((32, 32, 32), 7, "Discrete"), ((64, 64), 70, "Box"), ((128,), 7, "Discrete"),
                                                        ((32, 64), 7, "Discrete")])
def test_iteration_vs_oracle_on_shapes_outside_the_fused_kernels(hidden, obs_dim, action_type):
    """Three hidden layers (C3 / C5), observations wider than 64 (C5: 393), one layer, unequal widths: the tcgen05 fused
    rollout / update kernels decline these shapes and the layer-wise kernels take over inside the same entry points."""
    from harl_b200.runners import RUNNER_REGISTRY

    args, algo_args, env_args = small_config(action_type=action_type, hidden=hidden)
    env_args.update(obs_dim=obs_dim, share_obs_dim=obs_dim + 2)
    runner = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
    runner.warmup()
    runner.logger.init(2)
    check_iteration(runner)
    check_iteration(runner)
    runner.close()


@pytest.mark.parametrize("use_gae,ptl,vn", [(True, False, False), (False, True, True), (False, False, False)])
def test_iteration_return_branches(use_gae, ptl, vn):
    from harl_b200.runners import RUNNER_REGISTRY

    args, algo_args, env_args = small_config(use_gae=use_gae)
    algo_args["train"]["use_proper_time_limits"] = ptl
    algo_args["train"]["use_valuenorm"] = vn
    runner = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
    runner.warmup()
    runner.logger.init(1)
    check_iteration(runner)
    runner.close()


def test_haa2c_iteration():
    from harl_b200.runners import RUNNER_REGISTRY
    from oracle import algo as oa

    args, algo_args, env_args = small_config(algo="haa2c")
    runner = RUNNER_REGISTRY["haa2c"](args, algo_args, env_args)
    runner.warmup()
    runner.logger.init(1)
    # oracle without clipping = an unreachable clip range
    orig = oa.ppo_loss

    def no_clip(logp, old, adv, active, factor, ent, cfg):
        return orig(logp, old, adv, active, factor, ent, {**cfg, "clip_param": 1e30})

    oa.ppo_loss = no_clip
    try:
        check_iteration(runner)
    finally:
        oa.ppo_loss = orig
    runner.close()


@pytest.mark.parametrize("action_type,state_type,share", [("Discrete", "EP", False), ("Discrete", "EP", True), ("Box", "FP", True),
                                                         ("Discrete", "FP", False)])
def test_mappo_iteration_vs_oracle(action_type, state_type, share):
    """OnPolicyMARunner (reference on_policy_ma_runner.py:10-60) incl. MAPPO.share_param_train (mappo.py:149-222)."""
    from harl_b200.runners import RUNNER_REGISTRY

    args, algo_args, env_args = small_config(algo="mappo", action_type=action_type, state_type=state_type)
    algo_args["algo"]["share_param"] = share
    runner = RUNNER_REGISTRY["mappo"](args, algo_args, env_args)
    runner.warmup()
    runner.logger.init(2)
    check_iteration(runner)
    check_iteration(runner)
    runner.close()


def _replay_perms(monkeypatch, g):
    """torch.randperm replaced by the permutations the unmodified reference drew while the golden was recorded
    (tests/golden/make_golden.py PermRecorder), in the same call order."""
    it = iter([g[f"perm{i}"] for i in range(int(g["n_perms"]))])

    def fake(n, *a, **k):
        p = next(it)
        assert len(p) == n, (len(p), n)
        return torch.from_numpy(np.ascontiguousarray(p)).long()

    monkeypatch.setattr(torch, "randperm", fake)
    return it


def _load_runner_from_golden(g, cfg, m, actor_cls=None, runner_cls=None, share=False):
    """A runner shell (no env / dirs) holding the golden buffers and weights."""
    from harl_b200.algorithms.actors.happo import HAPPO
    from harl_b200.common.buffers.on_policy_critic_buffer_fp import OnPolicyCriticBufferFP
    from harl_b200.algorithms.critics.v_critic import VCritic
    from harl_b200.common.buffers.on_policy_actor_buffer import OnPolicyActorBuffer
    from harl_b200.common.buffers.on_policy_critic_buffer_ep import OnPolicyCriticBufferEP
    from harl_b200.common.valuenorm import ValueNorm
    from harl_b200.envs.spaces import Box, Discrete
    from harl_b200.runners.on_policy_ha_runner import OnPolicyHARunner

    dev = torch.device("cuda:0")
    A = m["A"]
    act_space = Discrete(m["act_dim"]) if m["head"] == "Discrete" else Box(shape=(m["act_dim"],))
    r = object.__new__(runner_cls or OnPolicyHARunner)
    r.share_param, r.world = share, 1
    r.overlap_critic_update = False   # the reference draws the actors' permutations before the critic's
    r.algo_args = {"train": cfg, "algo": cfg, "model": cfg}
    r.device, r.n_local, r.num_agents, r.state_type = dev, cfg["n_rollout_threads"], A, m["state_type"]
    r.fixed_order, r.action_aggregation = True, cfg["action_aggregation"]
    r.actor, r.actor_buffer = [], []
    for a in range(A):
        if share and a > 0:
            r.actor.append(r.actor[0])
        else:
            ac = (actor_cls or HAPPO)(cfg, Box(shape=(m["od"],)), act_space, device=dev)
            ac.actor.load_state_dict(U.params_of(g, f"actor{a}/"))
            r.actor.append(ac)
        b = OnPolicyActorBuffer(cfg, Box(shape=(m["od"],)), act_space, device=dev)
        for k in ("obs", "actions", "action_log_probs", "masks", "active_masks") + (("rnn_states",) if b.recurrent else ()):
            getattr(b, k).copy_(torch.from_numpy(g[f"a{a}.{k}"]))
        if b.available_actions is not None:
            b.available_actions.copy_(torch.from_numpy(g[f"a{a}.available_actions"]))
        r.actor_buffer.append(b)
    r.critic = VCritic(cfg, Box(shape=(m["sd"],)), device=dev)
    r.critic.critic.load_state_dict(U.params_of(g, "critic/"))
    if m["state_type"] == "FP":
        cb = OnPolicyCriticBufferFP(cfg, Box(shape=(m["sd"],)), A, device=dev)
    else:
        cb = OnPolicyCriticBufferEP(cfg, Box(shape=(m["sd"],)), device=dev)
    for k in ("share_obs", "value_preds", "returns", "rewards", "masks", "bad_masks") + (("rnn_states_critic",) if cb.recurrent else ()):
        getattr(cb, k).copy_(torch.from_numpy(g["c." + k]))
    r.critic_buffer = cb
    r.value_normalizer = ValueNorm(1, device=dev)
    r.value_normalizer.state.copy_(torch.from_numpy(g["vn_in"]))
    # advantages as the runner would have them after compute(): returns[:-1] - denorm(value_preds[:-1])
    den = r.value_normalizer.denormalize(cb.value_preds[:-1])
    cb.advantages.copy_(cb.returns[:-1] - torch.from_numpy(den).to(dev))
    return r


@pytest.mark.parametrize("name", ["ha_train_mlp_disc_EP", "ha_train_mlp_box_EP", "ha_train_mlp_disc_mb2_EP"])
def test_reference_ha_train_golden(name, monkeypatch):
    """The unmodified reference's OnPolicyHARunner.train() outputs, reproduced by the device path (two minibatches:
    on the reference's own recorded permutations)."""
    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    r = _load_runner_from_golden(g, cfg, m)
    if cfg["actor_num_mini_batch"] > 1 or cfg["critic_num_mini_batch"] > 1:
        left = _replay_perms(monkeypatch, g)
    infos, cinfo = r.train()
    if cfg["actor_num_mini_batch"] > 1 or cfg["critic_num_mini_batch"] > 1:
        assert next(left, None) is None   # every recorded permutation was consumed
    torch.cuda.synchronize()
    for a in range(m["A"]):
        np.testing.assert_allclose(r.actor_buffer[a].factor.cpu().numpy(), g[f"out.factor{a}"], rtol=3e-4, atol=3e-5)
        got = [infos[a][k] for k in ("policy_loss", "dist_entropy", "actor_grad_norm", "ratio")]
        np.testing.assert_allclose(got, g[f"out.info{a}"], rtol=3e-4, atol=3e-5)
        for k, v in r.actor[a].actor.state_dict().items():
            np.testing.assert_allclose(v.cpu().numpy(), g[f"out.actor{a}/" + k], rtol=0, atol=3e-5, err_msg=k)
    np.testing.assert_allclose([cinfo["value_loss"], cinfo["critic_grad_norm"]], g["out.cinfo"], rtol=3e-4)
    for k, v in r.critic.critic.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), g["out.critic/" + k], rtol=0, atol=3e-5, err_msg=k)
    np.testing.assert_allclose(r.value_normalizer.state.cpu().numpy(), g["out.vn"], rtol=1e-5)


@pytest.mark.parametrize("name", U.names("ma_train_"))
def test_reference_ma_train_golden(name, monkeypatch):
    """The unmodified reference's OnPolicyMARunner.train() (MAPPO; separate actors, shared parameters, shared
    parameters with two minibatches and FP state) reproduced by the device path."""
    from harl_b200.algorithms.actors.mappo import MAPPO
    from harl_b200.runners.on_policy_ma_runner import OnPolicyMARunner

    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    share = bool(int(g["share"][0]))
    r = _load_runner_from_golden(g, cfg, m, actor_cls=MAPPO, runner_cls=OnPolicyMARunner, share=share)
    multi = cfg["actor_num_mini_batch"] > 1 or cfg["critic_num_mini_batch"] > 1
    if multi:
        left = _replay_perms(monkeypatch, g)
    infos, cinfo = r.train()
    torch.cuda.synchronize()
    if multi:
        assert next(left, None) is None
    for a in range(m["A"]):
        got = [infos[a][k] for k in ("policy_loss", "dist_entropy", "actor_grad_norm", "ratio")]
        np.testing.assert_allclose(got, g[f"out.info{a}"], rtol=3e-4, atol=3e-5)
        for k, v in r.actor[a].actor.state_dict().items():
            np.testing.assert_allclose(v.cpu().numpy(), g[f"out.actor{a}/" + k], rtol=0, atol=3e-5, err_msg=k)
    np.testing.assert_allclose([cinfo["value_loss"], cinfo["critic_grad_norm"]], g["out.cinfo"], rtol=3e-4)
    for k, v in r.critic.critic.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), g["out.critic/" + k], rtol=0, atol=3e-5, err_msg=k)
    np.testing.assert_allclose(r.value_normalizer.state.cpu().numpy(), g["out.vn"], rtol=1e-5)


@pytest.mark.parametrize("action_type,state_type,simple", [("Discrete", "EP", True), ("Box", "FP", False), ("Discrete", "FP", False)])
def test_zero_copy_rollout_equals_generic_rollout(action_type, state_type, simple):
    """The lean rollout loop (hb_rollout_collect + env.step_into + insert kernel) must fill the buffers exactly
    like the reference-shaped collect / step / insert loop, and lead to the same update."""
    from harl_b200.runners import RUNNER_REGISTRY

    runners = []
    for fast in (True, False):
        args, algo_args, env_args = small_config(action_type=action_type, state_type=state_type)
        if simple:
            env_args.update(death_prob=0.0, terminate_prob=0.0, avail_prob=1.0)
        algo_args["algo"]["fixed_order"] = True
        algo_args["train"]["log_interval"] = 10**9  # keep the episode-return accumulators running
        r = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
        r.disable_fast_rollout = not fast
        r.warmup()
        r.logger.init(3)
        # iteration 1: identical weights and sampling streams -> the rollout must be bit-identical
        r.run_iteration(1, 3)
        torch.cuda.synchronize()
        assert bool(r._fast) == fast
        snap = {}
        for a in range(r.num_agents):
            b = r.actor_buffer[a]
            for k in ("obs", "actions", "action_log_probs", "masks", "active_masks", "available_actions"):
                if getattr(b, k) is not None:
                    snap[f"a{a}.{k}"] = getattr(b, k).clone()
        for k in ("share_obs", "value_preds", "returns", "rewards", "masks", "bad_masks"):
            snap["c." + k] = getattr(r.critic_buffer, k).clone()
        r.snap = snap
        # iteration 2 runs on weights that differ in the last bits (atomic summation order): compare loosely
        r.run_iteration(2, 3)
        torch.cuda.synchronize()
        runners.append(r)
    f, g = runners
    assert f.snap.keys() == g.snap.keys()
    from harl_b200 import _lib as L

    exact = L.lib.hb_get_gemm_impl() == 0  # the fused rollout kernel is FP32 FFMA; the generic path follows the GEMM mode
    for k in f.snap:
        if exact or k.split(".")[1] in ("obs", "masks", "active_masks", "available_actions", "share_obs", "rewards", "bad_masks"):
            assert torch.equal(f.snap[k], g.snap[k]), k
        else:  # a sampled discrete action can flip when a probability moves by 1e-7: compare all but a few entries
            a_, b_ = f.snap[k], g.snap[k]
            close = ((a_ - b_).abs() <= 1e-3 + 1e-3 * b_.abs()).float().mean().item()
            assert close > 0.99, (k, close)
    for a in range(f.num_agents):
        for (k, v), (_, w) in zip(f.actor[a].actor.state_dict().items(), g.actor[a].actor.state_dict().items()):
            np.testing.assert_allclose(v.cpu().numpy(), w.cpu().numpy(), rtol=0, atol=1e-4, err_msg=k)
    # logger bookkeeping: device accumulators of the fast path vs the per_step path
    fs, gs = f.logger.done_sum.cpu().numpy(), g.logger.done_sum.cpu().numpy()
    np.testing.assert_allclose(fs, gs, rtol=1e-5)
    for r in runners:
        r.close()


@pytest.mark.parametrize("action_type,state_type,simple", [("Discrete", "EP", True), ("Box", "FP", False), ("Discrete", "EP", False)])
def test_host_staged_rollout_equals_device_rollout(action_type, state_type, simple):
    """A host-resident env (pinned staging, H2D of every output, D2H of the actions) must fill the buffers
    bit-identically to the device-resident env, and the env must have seen the sampled actions on the host."""
    from harl_b200.runners import RUNNER_REGISTRY

    snaps, runners = [], []
    for host in (True, False):
        args, algo_args, env_args = small_config(action_type=action_type, state_type=state_type)
        env_args["host"] = host
        if simple:
            env_args.update(death_prob=0.0, terminate_prob=0.0, avail_prob=1.0)
        algo_args["algo"]["fixed_order"] = True
        r = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
        r.warmup()
        r.logger.init(2)
        r.run_iteration(1, 2)
        torch.cuda.synchronize()
        assert r._fast, "the staged host env must take the lean rollout loop"
        snap = {}
        for a in range(r.num_agents):
            b = r.actor_buffer[a]
            for k in ("obs", "actions", "action_log_probs", "masks", "active_masks", "available_actions"):
                if getattr(b, k) is not None:
                    snap[f"a{a}.{k}"] = getattr(b, k).clone()
        for k in ("share_obs", "value_preds", "returns", "rewards", "masks", "bad_masks"):
            snap["c." + k] = getattr(r.critic_buffer, k).clone()
        snaps.append(snap)
        runners.append(r)
    h, d = snaps
    for k in h:
        assert torch.equal(h[k], d[k]), k
    env = runners[0].envs
    T = runners[0].algo_args["train"]["episode_length"]
    assert env.h2d_bytes > 0 and env.d2h_bytes == T * sum(x.numel() * 4 for x in env.last_actions)
    last = runners[0].actor_buffer[0].actions[T - 1].cpu()
    assert torch.equal(env.last_actions[0], last)
    for r in runners:
        r.close()


@pytest.mark.parametrize("action_type,state_type,simple,recurrent", [
    ("Discrete", "EP", True, False), ("Box", "FP", False, False), ("Discrete", "EP", False, False),
    ("Discrete", "FP", False, True), ("Box", "EP", False, True)])
def test_cuda_graph_rollout_equals_eager_rollout(action_type, state_type, simple, recurrent):
    """The T-step rollout replayed from a CUDA graph (iterations >= 2) must fill the buffers bit-identically to the
    eager loop: same kernels, the Philox offsets advanced through the device counter.  Learning rates are zero so
    that the two runs see identical weights in every iteration."""
    from harl_b200.runners import RUNNER_REGISTRY

    snaps = []
    for graph in (True, False):
        args, algo_args, env_args = small_config(action_type=action_type, state_type=state_type, T=8)
        env_args["pool"] = 4
        if simple:
            env_args.update(death_prob=0.0, terminate_prob=0.0, avail_prob=1.0)
        algo_args["model"]["lr"] = 0.0
        algo_args["model"]["critic_lr"] = 0.0
        if recurrent:  # GRU actors and critic: per-net kernels inside the graph, hidden states carried across replays
            algo_args["model"].update(use_recurrent_policy=True, data_chunk_length=4)
        algo_args["train"]["use_linear_lr_decay"] = False
        algo_args["train"]["log_interval"] = 10**9
        algo_args["algo"]["fixed_order"] = True
        r = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
        r.use_cuda_graph_rollout = graph
        r.warmup()
        r.logger.init(4)
        per_iter = []
        for it in range(1, 5):
            r.run_iteration(it, 4)
            torch.cuda.synchronize()
            snap = {}
            for a in range(r.num_agents):
                b = r.actor_buffer[a]
                for k in ("obs", "actions", "action_log_probs", "masks", "active_masks", "available_actions") + (("rnn_states",) if recurrent else ()):
                    if getattr(b, k) is not None:
                        snap[f"a{a}.{k}"] = getattr(b, k).clone()
            for k in ("share_obs", "value_preds", "returns", "rewards", "masks", "bad_masks") + (("rnn_states_critic",) if recurrent else ()):
                snap["c." + k] = getattr(r.critic_buffer, k).clone()
            per_iter.append(snap)
        assert (r._fast.get("graph") is not None) == graph
        snaps.append(per_iter)
        done_sum = r.logger.done_sum.cpu().numpy()
        snaps.append(done_sum)
        r.close()
    g_iters, g_done, e_iters, e_done = snaps
    for it, (g, e) in enumerate(zip(g_iters, e_iters)):
        for k in g:
            assert torch.equal(g[k], e[k]), (it, k)
    # consecutive iterations must not repeat the same samples (the offset counter advances inside the graph)
    assert not torch.equal(g_iters[2]["a0.actions"], g_iters[3]["a0.actions"])
    np.testing.assert_allclose(g_done, e_done, rtol=1e-6)
