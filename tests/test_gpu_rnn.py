"""GPU parity of the recurrent (GRU) path, through the C-ABI: RNNLayer forward in both modes, BPTT gradients, the
recurrent minibatch index maps, and the reference's own OnPolicyHARunner.train() goldens with GRU policies.

Tolerances: log-probs / values / hidden states 3e-5 abs; gradients 3e-4 of the tensor max (sums over T steps of
BPTT); weights after the update 5e-5 abs; factors 5e-4 rel; train-info scalars 5e-4.
"""
import os

import numpy as np
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
GRU_CFG = {
    "gru_disc": (dict(use_recurrent_policy=True, hidden_sizes=[16, 16]), "Discrete"),
    "gru_box": (dict(use_recurrent_policy=True, hidden_sizes=[16], recurrent_n=2), "Box"),
}


def _cu(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).to(DEV).contiguous()


def _net(cfg, in_dim, head, out_dim, params):
    from harl_b200 import _lib as L
    from harl_b200.nets import DeviceNet

    hid = {"Discrete": L.HEAD_DISCRETE, "Box": L.HEAD_BOX, "value": L.HEAD_VALUE}[head]
    net = DeviceNet(cfg, in_dim, hid, out_dim, torch.device(DEV), init=False)
    net.load_state_dict(params)
    return net


@pytest.mark.parametrize("tag", sorted(GRU_CFG))
def test_gru_policy_forward_golden(tag):
    """StochasticPolicy / VNet with a GRU vs the unmodified reference: sequence mode (T steps from [N, R, h]
    states, rnn.py:33-78), row mode (one step per row, rnn.py:24-32) and the new hidden states."""
    from harl_b200.nets import DeviceNet

    over, head = GRU_CFG[tag]
    cfg = U.base_args(**over)
    g = U.load(f"policy_{tag}")
    od = g["obs"].shape[1]
    out_dim = g["avail"].shape[1] if head == "Discrete" else g["actions"].shape[1]
    net = _net(cfg, od, head, out_dim, U.params_of(g, "actor/"))
    critic = _net(cfg, g["cobs"].shape[1], "value", 1, U.params_of(g, "critic/"))
    obs, acts, cobs = _cu(g["obs"]), _cu(g["actions"]), _cu(g["cobs"])
    avail = _cu(g["avail"]) if "avail" in g else None
    masks = _cu(g["masks"]).reshape(-1)
    B = obs.shape[0]
    for mode in ("seq", "row"):
        hx = _cu(g[f"{mode}.hxs"])
        n_seq = hx.shape[0]
        hx2 = hx.reshape(n_seq, -1).contiguous()
        logp = torch.zeros(B, net.act_width, device=DEV)
        net.evaluate(DeviceNet.actor_batch(obs, acts, avail=avail, rnn_states=hx2, masks=masks, seq_len=B // n_seq),
                     logp_out=logp)
        np.testing.assert_allclose(logp.cpu().numpy(), g[f"{mode}.logp"], rtol=1e-5, atol=3e-5, err_msg=mode)
    # one rollout step per row: values + new hidden states of the critic, deterministic actions of the actor
    hx = _cu(g["row.hxs"])
    v = torch.zeros(B, 1, device=DEV)
    hc = torch.zeros_like(hx)
    critic.values(cobs, v, hx.reshape(B, -1).contiguous(), masks, hc.reshape(B, -1))
    np.testing.assert_allclose(v.cpu().numpy(), g["row.values"], rtol=1e-5, atol=3e-5)
    np.testing.assert_allclose(hc.cpu().numpy(), g["row.hxs_critic_out"], rtol=1e-5, atol=3e-5)
    a = torch.zeros(B, net.act_width, device=DEV)
    lp = torch.zeros(B, net.act_width, device=DEV)
    hn = torch.zeros_like(hx)
    net.act(obs, avail, True, 0, 0, a, lp, hx.reshape(B, -1).contiguous(), masks, hn.reshape(B, -1))
    if head == "Discrete":
        assert np.array_equal(a.cpu().numpy(), g["det_action"])
    else:
        np.testing.assert_allclose(a.cpu().numpy(), g["det_action"], rtol=1e-5, atol=3e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), g["det_logp"], rtol=1e-5, atol=3e-5)
    np.testing.assert_allclose(hn.cpu().numpy(), g["det_hxs"], rtol=1e-5, atol=3e-5)


@pytest.mark.parametrize("mode,head", [("chunk", "Discrete"), ("chunk", "Box"), ("naive", "Discrete")])
@pytest.mark.parametrize("recurrent_n", [1, 2])
def test_gru_actor_and_critic_gradients_vs_oracle(mode, head, recurrent_n):
    """hb_ppo_actor_grad / hb_value_grad on recurrent minibatches (chunk and whole-trajectory index maps) vs the
    oracle's autograd through the same sequences."""
    from harl_b200 import _lib as L
    from harl_b200.common import seq_index
    from harl_b200.nets import DeviceNet
    from oracle import algo as oa
    from oracle import nets as on

    T, N, od, sd, h = 8, 6, 7, 9, 16
    cfg = U.base_args(hidden_sizes=[h, h], recurrent_n=recurrent_n, episode_length=T, n_rollout_threads=N,
                      use_recurrent_policy=mode == "chunk", use_naive_recurrent_policy=mode == "naive",
                      data_chunk_length=4, actor_num_mini_batch=2, critic_num_mini_batch=2)
    na = 5 if head == "Discrete" else 3
    rng = np.random.default_rng(7 + recurrent_n)
    torch.manual_seed(11)
    p = on.init_params(cfg, od, head, na)
    pc = on.init_params(cfg, sd, "value", 1)
    for q in (p, pc):
        for k in q:
            q[k] = (q[k] + 0.1 * torch.randn(q[k].shape)).requires_grad_(True)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    ad = 1 if head == "Discrete" else na
    buf = dict(obs=f(T + 1, N, od), rnn_states=f(T + 1, N, recurrent_n, h), masks=(rng.random((T + 1, N, 1)) > 0.25).astype(np.float32),
               active_masks=(rng.random((T + 1, N, 1)) > 0.2).astype(np.float32),
               actions=(rng.integers(0, na, (T, N, 1)).astype(np.float32) if head == "Discrete" else f(T, N, na)),
               action_log_probs=-np.abs(f(T, N, ad)) - 0.5,
               available_actions=None)
    if head == "Discrete":
        av = (rng.random((T + 1, N, na)) < 0.7).astype(np.float32)
        av[np.arange(T)[:, None], np.arange(N)[None, :], buf["actions"][..., 0].astype(int)] = 1.0
        buf["available_actions"] = av
    adv, factor = f(T, N, 1), (1 + 0.1 * f(T, N, 1)).astype(np.float32)
    cbuf = dict(share_obs=f(T + 1, N, sd), rnn_states_critic=f(T + 1, N, recurrent_n, h), value_preds=f(T + 1, N, 1),
                returns=f(T + 1, N, 1), masks=buf["masks"].copy())
    # ---- draw the permutations once and replay them on both sides
    perms = []
    orig = torch.randperm

    def rec(n, *a, **k):
        q = orig(n)
        perms.append(q.numpy().copy())
        return q

    torch.randperm = rec
    try:
        parts = list(seq_index.minibatches(T, N, 2, mode, cfg["data_chunk_length"], torch.device(DEV)))
    finally:
        torch.randperm = orig
    assert len(perms) == 1
    o_batches = list(oa.actor_minibatches(buf, adv, factor, cfg, lambda n: perms[0]))
    c_batches = list(oa.critic_minibatches(cbuf, cfg, lambda n: perms[0]))
    net = _net(cfg, od, head, na, {k: v.detach() for k, v in p.items()})
    cnet = _net(cfg, sd, "value", 1, {k: v.detach() for k, v in pc.items()})
    rows = T * N
    fl = lambda a: _cu(a.reshape(rows, *a.shape[2:]))
    d_rnn = _cu(buf["rnn_states"].reshape((T + 1) * N, -1))
    d_masks = _cu(buf["masks"].reshape((T + 1) * N))
    d_crnn = _cu(cbuf["rnn_states_critic"].reshape((T + 1) * N, -1))
    hyper = L.PPOHyper(0.2, 0.01, 1, 1, 1)
    vh = L.ValueHyper(0.2, 10.0, 1.0, 1, 1)
    for (idx, nrows, seq_len), ob_, cb_ in zip(parts, o_batches, c_batches):
        # oracle
        logp, ent, _, _ = on.actor_evaluate(p, cfg, head, ob_["obs"], ob_["rnn"], ob_["actions"], ob_["masks"],
                                            ob_.get("avail"), ob_["active"])
        pl, total, _ = oa.ppo_loss(logp, ob_["old_logp"], ob_["adv"], ob_["active"], ob_["factor"], ent, cfg)
        names = list(p.keys())
        gs = torch.autograd.grad(total, [p[k] for k in names], allow_unused=True)
        # device
        batch = DeviceNet.actor_batch(fl(buf["obs"][:-1]), fl(buf["actions"]), fl(buf["action_log_probs"]), fl(adv).reshape(-1),
                                      fl(factor).reshape(-1), fl(buf["active_masks"][:-1]).reshape(-1),
                                      None if buf["available_actions"] is None else fl(buf["available_actions"][:-1]),
                                      idx, nrows, rnn_states=d_rnn, masks=d_masks, seq_len=seq_len)
        norm3 = torch.zeros(3, dtype=torch.float64, device=DEV)
        norm3[2] = float(ob_["active"].sum())
        scal = torch.zeros(4, dtype=torch.float64, device=DEV)
        net.actor_grad(batch, hyper, norm3, scal)
        torch.cuda.synchronize()
        np.testing.assert_allclose(scal[0].item() / norm3[2].item(), float(pl), rtol=2e-4, atol=1e-6)
        got = {k: v.cpu().numpy() for k, v in net.views(net.grad).items()}
        for k, gi in zip(names, gs):
            ref = np.zeros_like(got[k]) if gi is None else gi.numpy()
            scale = max(np.abs(ref).max(), 1e-6)
            assert np.abs(got[k] - ref).max() <= 3e-4 * max(scale, 1e-3), f"actor grad {k}: {np.abs(got[k] - ref).max()} vs {scale}"
        # critic
        values, _ = on.critic_values(pc, cfg, cb_["share_obs"], cb_["rnn"], cb_["masks"])
        vl = oa.value_loss(values, cb_["value_preds"], cb_["returns"], cfg, None)
        cn = list(pc.keys())
        cgs = torch.autograd.grad(vl, [pc[k] for k in cn])
        cbatch = DeviceNet.critic_batch(fl(cbuf["share_obs"][:-1]), fl(cbuf["value_preds"][:-1]).reshape(-1),
                                        fl(cbuf["returns"][:-1]).reshape(-1), idx, nrows, d_crnn, d_masks, seq_len)
        cscal = torch.zeros(4, dtype=torch.float64, device=DEV)
        cnet.value_grad(cbatch, vh, None, 1.0 / nrows, cscal)
        torch.cuda.synchronize()
        np.testing.assert_allclose(cscal[0].item() / cscal[1].item(), float(vl), rtol=2e-4)
        cgot = {k: v.cpu().numpy() for k, v in cnet.views(cnet.grad).items()}
        for k, gi in zip(cn, cgs):
            scale = max(np.abs(gi.numpy()).max(), 1e-3)
            assert np.abs(cgot[k] - gi.numpy()).max() <= 3e-4 * scale, f"critic grad {k}"


@pytest.mark.parametrize("name", ["ha_train_gru_box_EP", "ha_train_naive_gru_disc_EP", "ha_train_gru_disc_FP"])
def test_reference_ha_train_gru_golden(name):
    """The unmodified reference's OnPolicyHARunner.train() with GRU policies, reproduced on the device (one
    minibatch per epoch, so the permutation the reference drew only reorders sums)."""
    from tests.test_gpu_iteration import _load_runner_from_golden

    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    r = _load_runner_from_golden(g, cfg, m)
    infos, cinfo = r.train()
    torch.cuda.synchronize()
    for a in range(m["A"]):
        np.testing.assert_allclose(r.actor_buffer[a].factor.cpu().numpy(), g[f"out.factor{a}"], rtol=5e-4, atol=5e-5)
        got = [infos[a][k] for k in ("policy_loss", "dist_entropy", "actor_grad_norm", "ratio")]
        np.testing.assert_allclose(got, g[f"out.info{a}"], rtol=5e-4, atol=5e-5)
        for k, v in r.actor[a].actor.state_dict().items():
            np.testing.assert_allclose(v.cpu().numpy(), g[f"out.actor{a}/" + k], rtol=0, atol=5e-5, err_msg=k)
    np.testing.assert_allclose([cinfo["value_loss"], cinfo["critic_grad_norm"]], g["out.cinfo"], rtol=5e-4)
    for k, v in r.critic.critic.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), g["out.critic/" + k], rtol=0, atol=5e-5, err_msg=k)
    np.testing.assert_allclose(r.value_normalizer.state.cpu().numpy(), g["out.vn"], rtol=1e-5)


@pytest.mark.parametrize("algo,over", [("happo", dict(use_recurrent_policy=True)),
                                       ("happo", dict(use_naive_recurrent_policy=True)),
                                       ("mappo", dict(use_recurrent_policy=True))])
@pytest.mark.parametrize("state_type", ["EP", "FP"])
def test_recurrent_iteration_vs_oracle(algo, over, state_type):
    """Whole iterations through the public runner with GRU actors and critic (rollout with hidden-state
    propagation and resets, GAE, recurrent updates), replayed by the oracle; two iterations so the slot T -> 0
    carry-over of the hidden states is covered."""
    from harl_b200.runners import RUNNER_REGISTRY
    from tests.smoke_check import check_iteration, small_config

    args, algo_args, env_args = small_config(algo=algo, state_type=state_type, T=12)
    algo_args["model"].update(data_chunk_length=4, **over)
    algo_args["algo"]["share_param"] = False
    runner = RUNNER_REGISTRY[algo](args, algo_args, env_args)
    runner.warmup()
    runner.logger.init(2)
    check_iteration(runner, tol_w=5e-5, tol_info=5e-4)
    check_iteration(runner, tol_w=5e-5, tol_info=5e-4)
    runner.close()


def test_share_param_with_recurrent_policy_fails_loudly():
    """The reference interleaves the agents' sequences in this combination (see MAPPO.share_param_train)."""
    from harl_b200.runners import RUNNER_REGISTRY
    from tests.smoke_check import small_config

    args, algo_args, env_args = small_config(algo="mappo", T=8)
    algo_args["model"].update(data_chunk_length=4, use_recurrent_policy=True)
    algo_args["algo"]["share_param"] = True
    with pytest.raises(NotImplementedError, match="share_param"):   # at construction, before the first rollout
        RUNNER_REGISTRY["mappo"](args, algo_args, env_args)


@pytest.mark.parametrize("state_type,over", [("EP", dict(use_recurrent_policy=True)), ("FP", dict(use_naive_recurrent_policy=True))])
def test_recurrent_zero_copy_rollout_equals_generic_rollout(state_type, over):
    """The lean rollout loop (one hb_rollout_collect per step for all agents + critic, env.step_into, insert kernel
    with the hidden-state reset) must fill the buffers -- hidden states included -- exactly like the reference-shaped
    collect / step / insert loop: both run the same per-net kernels with the same sampling streams."""
    from harl_b200.runners import RUNNER_REGISTRY
    from tests.smoke_check import small_config

    runners = []
    for fast in (True, False):
        args, algo_args, env_args = small_config(state_type=state_type, T=12)
        algo_args["model"].update(data_chunk_length=4, **over)
        algo_args["algo"]["fixed_order"] = True
        algo_args["train"]["log_interval"] = 10**9
        r = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
        r.disable_fast_rollout = not fast
        r.warmup()
        r.logger.init(2)
        r.run_iteration(1, 2)
        torch.cuda.synchronize()
        assert bool(r._fast) == fast
        snap = {}
        for a in range(r.num_agents):
            b = r.actor_buffer[a]
            for k in ("obs", "rnn_states", "actions", "action_log_probs", "masks", "active_masks", "available_actions"):
                if getattr(b, k) is not None:
                    snap[f"a{a}.{k}"] = getattr(b, k).clone()
        for k in ("share_obs", "rnn_states_critic", "value_preds", "returns", "rewards", "masks", "bad_masks"):
            snap["c." + k] = getattr(r.critic_buffer, k).clone()
        r.snap = snap
        runners.append(r)
    f, g = runners
    assert f.snap.keys() == g.snap.keys()
    for k in f.snap:
        assert torch.equal(f.snap[k], g.snap[k]), k
    for a in range(f.num_agents):
        for (k, v), (_, w) in zip(f.actor[a].actor.state_dict().items(), g.actor[a].actor.state_dict().items()):
            np.testing.assert_allclose(v.cpu().numpy(), w.cpu().numpy(), rtol=0, atol=1e-4, err_msg=k)
    for r in runners:
        r.close()


@pytest.mark.parametrize("recurrent_n", [1, 2])
def test_persistent_recurrence_equals_per_step_kernels(recurrent_n):
    """hb_set_rnn_impl(1) must reproduce the launch-per-step path (same accumulation order: bit-identical states)."""
    from harl_b200 import _lib as L
    from harl_b200.nets import DeviceNet
    from oracle import nets as on

    T, N, od, h, na = 12, 40, 9, 64, 6
    cfg = U.base_args(hidden_sizes=[h, h], recurrent_n=recurrent_n, use_recurrent_policy=True)
    torch.manual_seed(3)
    p = on.init_params(cfg, od, "Discrete", na)
    net = _net(cfg, od, "Discrete", na, {k: v + 0.1 * torch.randn(v.shape) for k, v in p.items()})
    g = torch.Generator().manual_seed(4)
    B = T * N
    obs = torch.randn(B, od, generator=g).to(DEV)
    acts = torch.randint(0, na, (B, 1), generator=g).float().to(DEV)
    masks = (torch.rand(B, generator=g) > 0.2).float().to(DEV)
    hx = torch.randn(N, recurrent_n * h, generator=g).to(DEV)
    outs = []
    for impl in (0, 1):
        L.call("hb_set_rnn_impl", impl)
        try:
            lp = torch.zeros(B, 1, device=DEV)
            net.evaluate(DeviceNet.actor_batch(obs, acts, rnn_states=hx, masks=masks, seq_len=T), logp_out=lp)
            a = torch.zeros(N, 1, device=DEV)
            l1 = torch.zeros(N, 1, device=DEV)
            hn = torch.zeros(N, recurrent_n * h, device=DEV)
            net.act(obs[:N].contiguous(), None, True, 0, 0, a, l1, hx, masks[:N].contiguous(), hn)
            torch.cuda.synchronize()
            outs.append((lp.clone(), hn.clone()))
        finally:
            L.call("hb_set_rnn_impl", 0)
    np.testing.assert_allclose(outs[1][0].cpu().numpy(), outs[0][0].cpu().numpy(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(outs[1][1].cpu().numpy(), outs[0][1].cpu().numpy(), rtol=0, atol=1e-6)
