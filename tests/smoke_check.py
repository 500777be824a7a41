"""One small HAPPO iteration on cuda:0 through the public runner, checked against the CPU oracle.

Used by ``__graft_entry__.smoke()`` and by tests/test_gpu_iteration.py.  Teacher-forced: the
rollout is produced by the CUDA path (sampling cannot be RNG-matched), then the oracle recomputes
masks, returns, advantages and the whole sequential-agent update from the same buffers and the
same initial weights, and everything is compared.
"""
import copy
import tempfile

import numpy as np
import torch


def small_config(algo="happo", action_type="Discrete", state_type="EP", n=12, T=10, hidden=(32, 32), A=3, **algo_over):
    from harl_b200.utils.configs_tools import get_defaults_yaml_args

    algo_args, _ = get_defaults_yaml_args(algo, "synthetic")
    algo_args["train"].update(n_rollout_threads=n, episode_length=T, num_env_steps=n * T * 2, log_interval=1,
                              eval_interval=10**9)
    algo_args["eval"]["use_eval"] = False
    algo_args["model"]["hidden_sizes"] = list(hidden)
    algo_args["algo"].update(fixed_order=False, **algo_over)
    algo_args["logger"]["log_dir"] = tempfile.mkdtemp(prefix="harl_b200_")
    env_args = dict(task="unit", n_agents=A, obs_dim=7, share_obs_dim=9, action_type=action_type,
                    action_dim=5 if action_type == "Discrete" else 2, state_type=state_type, episode_limit=4,
                    death_prob=0.1, terminate_prob=0.05, avail_prob=0.7)
    args = dict(algo=algo, env="synthetic", exp_name="smoke", load_config="")
    return args, algo_args, env_args


def snapshot(runner):
    """Host copies of everything the update reads (buffers, weights, ValueNorm) in oracle form."""
    n = lambda t: t.detach().cpu().numpy().copy()
    abufs = []
    for b in runner.actor_buffer:
        d = dict(obs=n(b.obs), rnn_states=n(b.rnn_states), actions=n(b.actions), action_log_probs=n(b.action_log_probs),
                 masks=n(b.masks), active_masks=n(b.active_masks),
                 available_actions=None if b.available_actions is None else n(b.available_actions))
        abufs.append(d)
    cb = runner.critic_buffer
    cbuf = {k: n(getattr(cb, k)) for k in ("share_obs", "rnn_states_critic", "value_preds", "returns", "rewards", "masks",
                                           "bad_masks")}
    def net_state(net):  # weights + Adam moments + step count
        cp = lambda flat: {k: v.cpu().clone() for k, v in net.views(flat).items()}
        return dict(p=cp(net.params), m=cp(net.exp_avg), v=cp(net.exp_avg_sq), t=net.adam_steps)

    actors = [net_state(a.actor) for a in runner.actor]
    critic = net_state(runner.critic.critic)
    vn = None if runner.value_normalizer is None else n(runner.value_normalizer.state)
    return abufs, cbuf, actors, critic, vn


def oracle_iteration(runner, snap, agent_order):
    from oracle import algo as oa
    from oracle import buffers as ob

    abufs, cbuf, actors, critic, vn_state = copy.deepcopy(snap)
    cfg = {**runner.algo_args["model"], **runner.algo_args["algo"], **runner.algo_args["train"]}
    cfg.setdefault("ppo_epoch", cfg.get("a2c_epoch"))
    vn = None
    if vn_state is not None:
        vn = ob.ValueNormState()
        vn.running_mean, vn.running_mean_sq, vn.debiasing_term = (np.float32(x) for x in vn_state)
    heads = [sp.__class__.__name__ for sp in runner.envs.action_space]
    def oracle_net(st, lr):
        p = {k: v.clone().requires_grad_(True) for k, v in st["p"].items()}
        opt = oa.Adam(p, lr, cfg["opti_eps"], cfg["weight_decay"])
        opt.m, opt.v, opt.t = {k: v.clone() for k, v in st["m"].items()}, {k: v.clone() for k, v in st["v"].items()}, st["t"]
        return p, opt

    o_actors = [oracle_net(st, cfg["lr"]) for st in actors]
    o_critic = oracle_net(critic, cfg["critic_lr"])
    ident = lambda m: np.arange(m)
    if runner.__class__.__name__ == "OnPolicyMARunner":
        if runner.share_param:   # one shared network: every list entry is the same (params, Adam) pair
            o_actors = [o_actors[0]] * len(o_actors)
        infos, cinfo = oa.ma_train(o_actors, o_critic, cfg, heads, abufs, cbuf, vn, runner.state_type, runner.share_param, ident)
        return infos, cinfo, None, o_actors, o_critic, vn
    infos, cinfo, factors, _ = oa.ha_train(o_actors, o_critic, cfg, heads, abufs, cbuf, vn, runner.state_type,
                                           agent_order, ident)
    return infos, cinfo, factors, o_actors, o_critic, vn


def check_iteration(runner, tol_w=3e-5, tol_info=2e-4):
    """Run rollout + compute + train on the device and compare every product with the oracle."""
    from oracle import buffers as ob

    T = runner.algo_args["train"]["episode_length"]
    runner.logger.episode_init(1)
    dones_log, bad_log = [], []
    for step in range(T):
        values, actions, logp, rnn, rnn_c = runner.collect(step)
        obs, share_obs, rewards, dones, infos, avail = runner.envs.step(actions)
        dones_log.append(dones.cpu().numpy().copy())
        bad_log.append(runner.envs.last_bad_transition.cpu().numpy().copy())
        runner.insert((obs, share_obs, rewards, dones, infos, avail, values, actions, logp, rnn, rnn_c))
    runner.compute()
    torch.cuda.synchronize()
    snap = snapshot(runner)
    abufs, cbuf, _, _, vn_state = snap
    # ---- masks: bit-exact vs the oracle's restatement of the reference insert()
    for t in range(T):
        masks, active, bad, _ = ob.derive_masks(dones_log[t], bad_log[t], runner.state_type)
        for a in range(runner.num_agents):
            assert np.array_equal(abufs[a]["masks"][t + 1], masks[:, a]), "masks differ"
            assert np.array_equal(abufs[a]["active_masks"][t + 1], active[:, a]), "active_masks differ"
        assert np.array_equal(cbuf["masks"][t + 1], masks[:, 0] if runner.state_type == "EP" else masks)
        assert np.array_equal(cbuf["bad_masks"][t + 1], bad)
    # ---- recurrent nets: every stored hidden state is one oracle GRU step from the previous slot (zero after a reset)
    if runner.actor_buffer[0].recurrent:
        from oracle import nets as on

        cfg_m = {**runner.algo_args["model"], **runner.algo_args["algo"]}
        _, _, actors_s, critic_s, _ = snap
        tt = lambda x: torch.from_numpy(np.ascontiguousarray(x))
        with torch.no_grad():
            for t in range(T):
                for a in range(runner.num_agents):
                    b = abufs[a]
                    _, hx = on.features(actors_s[a]["p"], cfg_m, tt(b["obs"][t]), tt(b["rnn_states"][t]), tt(b["masks"][t]))
                    want = hx.numpy() * b["masks"][t + 1][:, :, None]
                    np.testing.assert_allclose(b["rnn_states"][t + 1], want, rtol=0, atol=3e-5,
                                               err_msg=f"actor {a} hidden state at slot {t + 1}")
                so, rc, mk = cbuf["share_obs"][t], cbuf["rnn_states_critic"][t], cbuf["masks"][t]
                sd_, (R_, h_) = so.shape[-1], rc.shape[-2:]
                _, hx = on.features(critic_s["p"], cfg_m, tt(so.reshape(-1, sd_)), tt(rc.reshape(-1, R_, h_)), tt(mk.reshape(-1, 1)))
                want = hx.numpy().reshape(rc.shape) * cbuf["masks"][t + 1][..., None]
                np.testing.assert_allclose(cbuf["rnn_states_critic"][t + 1], want, rtol=0, atol=3e-5,
                                           err_msg=f"critic hidden state at slot {t + 1}")
    # ---- returns / advantages: bit-exact
    vn = None
    if vn_state is not None:
        vn = ob.ValueNormState()
        vn.running_mean, vn.running_mean_sq, vn.debiasing_term = (np.float32(x) for x in vn_state)
    algo = runner.algo_args["algo"]
    # the bootstrap value sits in value_preds[-1] (GAE branches) or returns[-1] (plain returns)
    next_value = cbuf["value_preds"][-1] if algo["use_gae"] else cbuf["returns"][-1]
    ret, _ = ob.compute_returns(cbuf["rewards"], cbuf["value_preds"], cbuf["masks"], cbuf["bad_masks"],
                                next_value, algo["gamma"], algo["gae_lambda"], algo["use_gae"],
                                runner.algo_args["train"]["use_proper_time_limits"], vn)
    assert np.array_equal(ret[:-1], cbuf["returns"][:-1]), "returns differ from the oracle"
    adv = ob.advantages(ret, cbuf["value_preds"], vn)
    assert np.array_equal(adv, runner.critic_buffer.advantages.cpu().numpy()), "advantages differ"
    # ---- the sequential-agent update
    runner.prep_training()
    infos, cinfo = runner.train()
    torch.cuda.synchronize()
    order = [int(a) for a in runner.last_agent_order]
    o_infos, o_cinfo, o_factors, o_actors, o_critic, o_vn = oracle_iteration(runner, snap, order)
    errs = []

    def cmp(got, want, msg, **kw):
        try:
            np.testing.assert_allclose(got, want, err_msg=msg, **kw)
        except AssertionError as e:
            lines = str(e).strip().splitlines()
            errs.append(f"{msg}: " + " | ".join(l.strip() for l in lines[3:6]))

    for a in range(runner.num_agents):
        if o_factors is not None:
            cmp(runner.actor_buffer[a].factor.cpu().numpy(), o_factors[a], f"factor[{a}]", rtol=3e-4, atol=3e-5)
        for k in ("policy_loss", "dist_entropy", "actor_grad_norm", "ratio"):
            cmp(infos[a][k], o_infos[a][k], f"{k}[{a}]", rtol=tol_info, atol=tol_info)
        for k, v in runner.actor[a].actor.state_dict().items():
            cmp(v.cpu().numpy(), o_actors[a][0][k].detach().numpy(), f"actor{a}/{k}", rtol=0, atol=tol_w)
    cmp([cinfo["value_loss"], cinfo["critic_grad_norm"]], [o_cinfo["value_loss"], o_cinfo["critic_grad_norm"]],
        "critic info", rtol=tol_info)
    for k, v in runner.critic.critic.state_dict().items():
        cmp(v.cpu().numpy(), o_critic[0][k].detach().numpy(), f"critic/{k}", rtol=0, atol=tol_w)
    if o_vn is not None:
        cmp(runner.value_normalizer.state.cpu().numpy(), [o_vn.running_mean, o_vn.running_mean_sq, o_vn.debiasing_term],
            "valuenorm state", rtol=1e-5)
    assert not errs, f"agent order {order}; {len(errs)} mismatches:\n" + "\n".join(errs)
    runner.after_update()
    return infos, cinfo


def run_smoke():
    from harl_b200.runners import RUNNER_REGISTRY

    args, algo_args, env_args = small_config()
    runner = RUNNER_REGISTRY["happo"](args, algo_args, env_args)
    runner.warmup()
    runner.logger.init(1)
    infos, cinfo = check_iteration(runner)
    runner.close()
    print("smoke ok:", {k: round(v, 5) for k, v in infos[0].items()}, {k: round(v, 5) for k, v in cinfo.items()})
