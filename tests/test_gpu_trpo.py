"""GPU parity of the trust-region (HATRPO) path, through the C-ABI, against golden vectors produced by the
unmodified reference (tests/golden/hatrpo_*.npz) and against the CPU oracle (oracle/trpo.py).

Tolerances (the reference differentiates the KL twice in fp32; the device evaluates the same operator in
Gauss-Newton form): surrogate gradient 2e-4 of the tensor max; Fisher-vector product 5e-4 of the vector max;
conjugate-gradient direction 5e-3 of its max (10 fp32 CG steps amplify rounding); parameters after the
line-search step 2e-4 abs (steps are ~1e-2); scalars 3e-3 rel.
"""
import numpy as np
import pytest
import torch

from tests import util as U

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
# the synthetic-SMAC-width goldens run from tests/test_gpu_zz_wide_heads.py (last in the suite: they cover a dispatch fix
# made after the round's GPU budget was spent and have not run on a GPU yet)
PARTS = [n for n in U.names("hatrpo_parts_") if "12_h64" not in n]
WIDE_PARTS = [n for n in U.names("hatrpo_parts_") if "12_h64" in n]


@pytest.fixture(params=["3xtf32", "fp32"])
def gemm_impl(request):
    from harl_b200 import _lib as L

    prev = L.lib.hb_get_gemm_impl()
    L.lib.hb_set_gemm_impl(L.GEMM_IMPLS[request.param])
    yield request.param
    L.lib.hb_set_gemm_impl(prev)


def _actor(g, cfg, m):
    from harl_b200.algorithms.actors.hatrpo import HATRPO
    from harl_b200.envs.spaces import Box, Discrete

    act_space = Discrete(m["act_dim"]) if m["head"] == "Discrete" else Box(shape=(m["act_dim"],))
    ac = HATRPO(cfg, Box(shape=(m["od"],)), act_space, device=torch.device(DEV))
    ac.actor.load_state_dict(U.params_of(g, "actor0/"))
    return ac


def _sample(g, cfg):
    """The reference's single minibatch with the identity permutation (feed-forward: time-major flatten of the
    buffer; recurrent: every chunk in order, step-major, with the hidden state stored at each chunk start)."""
    from oracle import algo as oa

    buf = U.sub(g, "a0.")
    buf.setdefault("available_actions", None)
    b = next(oa.actor_minibatches(buf, g["adv"], g["factor"], cfg, lambda n: np.arange(n)))
    n = lambda k: None if b.get(k) is None else b[k].numpy()
    return (n("obs"), n("rnn"), n("actions"), n("masks"), n("active"), n("old_logp"), n("adv"), n("avail"), n("factor"))


def _flat(net, g, prefix):
    flat = torch.zeros(net.total, dtype=torch.float32)
    for k, v in net.views(flat).items():
        v.copy_(torch.from_numpy(g[prefix + k]).reshape(v.shape))
    return flat.to(DEV)


def _named(net, flat):
    return {k: v.cpu().numpy() for k, v in net.views(flat).items()}


def _max_rel(net, flat, g, prefix):
    got = _named(net, flat)
    scale = max(np.abs(g[prefix + k]).max() for k in got)
    return max(np.abs(got[k] - g[prefix + k].reshape(got[k].shape)).max() for k in got) / scale


def _device_batch(ac, g, cfg):
    from harl_b200.algorithms.actors.on_policy_base import to_device
    from harl_b200.nets import DeviceNet

    obs, rnn, actions, masks, active, old_lp, adv, avail, factor = _sample(g, cfg)
    d = ac.device
    t = [to_device(x, d) for x in (obs, actions, old_lp, adv, factor, active, avail)]
    kw = {}
    if ac.recurrent:
        r = to_device(rnn, d)
        kw = dict(rnn_states=r.reshape(r.shape[0], -1), masks=to_device(masks, d).reshape(-1), seq_len=obs.shape[0] // r.shape[0])
    batch = DeviceNet.actor_batch(t[0], t[1], t[2], t[3].reshape(-1), t[4].reshape(-1), t[5].reshape(-1), t[6], **kw)
    norm = float(active.sum()) if cfg["use_policy_active_masks"] else float(obs.shape[0])
    return batch, norm, float(obs.shape[0])


@pytest.mark.parametrize("name", PARTS)
def test_surrogate_gradient_and_fvp_vs_reference(name, gemm_impl):
    from harl_b200 import _lib as L

    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    ac = _actor(g, cfg, m)
    net = ac.actor
    batch, norm, rows = _device_batch(ac, g, cfg)
    norm3 = torch.zeros(3, dtype=torch.float64, device=DEV)
    norm3[2] = norm
    scal = torch.zeros(4, dtype=torch.float64, device=DEV)
    net.actor_grad(batch, ac._hyper(), norm3, scal)
    torch.cuda.synchronize()
    np.testing.assert_allclose(-scal[0].item() / norm, g["loss"][0], rtol=1e-4, atol=1e-6)
    assert _max_rel(net, -net.grad, g, "loss_grad/") <= 2e-4
    # one Fisher-vector product on the reference's random vector
    vec = _flat(net, g, "vec/")
    old_dist = torch.empty(batch.rows, net.out_dim, dtype=torch.float32, device=DEV)
    net.trpo_old_dist(batch, old_dist)
    out = torch.empty(net.total, dtype=torch.float32, device=DEV)
    net.trpo_fvp(batch, old_dist, vec, 1.0 / rows, out)
    net.trpo_fvp_finish(vec, out, 0.1)
    torch.cuda.synchronize()
    assert _max_rel(net, out, g, "fvp/") <= 5e-4


@pytest.mark.parametrize("name", PARTS)
def test_update_vs_reference(name, gemm_impl):
    """HATRPO.update (CG direction, step scaling, backtracking line search incl. rejected trials) vs the reference."""
    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    ac = _actor(g, cfg, m)
    kl, improve, expected, ent, ratio = ac.update(_sample(g, cfg))
    torch.cuda.synchronize()
    # 2-layer GRU: ill-conditioned damped Fisher, two CPU evaluation orders already differ by 1.2 % in the CG direction
    # (tests/test_oracle_trpo.py::_loose)
    loose = "gru2" in name
    assert _max_rel(ac.actor, ac.last_update["step_dir"], g, "step_dir/") <= (5e-2 if loose else 5e-3)
    np.testing.assert_allclose([kl, improve, expected, ent, ratio], g["update_scalars"], rtol=5e-2 if loose else 3e-3, atol=2e-6)
    for k, v in ac.actor.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), g["out.actor0/" + k], rtol=0, atol=2e-3 if loose else 2e-4, err_msg=k)
    if name.endswith("box_reject"):
        assert not ac.last_update["accepted"] and ac.last_update["trials"] == cfg["ls_step"]
    if name.endswith("disc_backtrack"):
        assert ac.last_update["accepted"] and ac.last_update["trials"] == 3


def _load_runner(g, cfg, m):
    from harl_b200.algorithms.actors.hatrpo import HATRPO
    from tests.test_gpu_iteration import _load_runner_from_golden

    r = _load_runner_from_golden(g, cfg, m, actor_cls=HATRPO)
    return r


@pytest.mark.parametrize("name", U.names("hatrpo_train_"))
def test_reference_ha_train_hatrpo_golden(name):
    """The unmodified reference's OnPolicyHARunner.train() with HATRPO actors, reproduced on the device."""
    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    r = _load_runner(g, cfg, m)
    infos, cinfo = r.train()
    torch.cuda.synchronize()
    for a in range(m["A"]):
        np.testing.assert_allclose(r.actor_buffer[a].factor.cpu().numpy(), g[f"out.factor{a}"], rtol=3e-3, atol=2e-4)
        got = [infos[a][k] for k in ("kl", "dist_entropy", "loss_improve", "expected_improve", "ratio")]
        np.testing.assert_allclose(got, g[f"out.info{a}"], rtol=5e-3, atol=1e-5)
        for k, v in r.actor[a].actor.state_dict().items():
            np.testing.assert_allclose(v.cpu().numpy(), g[f"out.actor{a}/" + k], rtol=0, atol=3e-4, err_msg=k)
    np.testing.assert_allclose([cinfo["value_loss"], cinfo["critic_grad_norm"]], g["out.cinfo"], rtol=3e-4)


@pytest.mark.parametrize("state_type,model_over", [("EP", {}), ("FP", dict(use_recurrent_policy=True, data_chunk_length=4)),
                                                   ("EP", dict(use_naive_recurrent_policy=True))])
def test_hatrpo_iteration_through_runner(state_type, model_over):
    """--algo hatrpo through the public runner: rollout -> GAE -> sequential trust-region updates -> critic; the
    update of every agent is replayed by the oracle from the same buffers and initial weights.  The FP / GRU case is
    BASELINE.json's configs[3] in miniature (HATRPO, recurrent policies, per-agent critic inputs)."""
    import copy

    from harl_b200.runners import RUNNER_REGISTRY
    from oracle import algo as oa
    from oracle import buffers as ob
    from oracle import trpo as ot
    from tests.smoke_check import small_config, snapshot

    args, algo_args, env_args = small_config(algo="hatrpo", n=16, T=12, state_type=state_type)
    algo_args["model"].update(model_over)
    runner = RUNNER_REGISTRY["hatrpo"](args, algo_args, env_args)
    runner.warmup()
    runner.logger.init(1)
    runner.logger.episode_init(1)
    T = algo_args["train"]["episode_length"]
    for step in range(T):
        values, actions, logp, rnn, rnn_c = runner.collect(step)
        obs, share_obs, rewards, dones, infos, avail = runner.envs.step(actions)
        runner.insert((obs, share_obs, rewards, dones, infos, avail, values, actions, logp, rnn, rnn_c))
    runner.compute()
    torch.cuda.synchronize()
    abufs, cbuf, actors, critic, vn_state = copy.deepcopy(snapshot(runner))
    runner.prep_training()
    infos, cinfo = runner.train()
    torch.cuda.synchronize()
    order = [int(a) for a in runner.last_agent_order]
    cfg = {**algo_args["model"], **algo_args["algo"], **algo_args["train"]}
    vn = ob.ValueNormState()
    vn.running_mean, vn.running_mean_sq, vn.debiasing_term = (np.float32(x) for x in vn_state)
    heads = [sp.__class__.__name__ for sp in runner.envs.action_space]
    o_actors = [{k: v.clone().requires_grad_(True) for k, v in st["p"].items()} for st in actors]
    pc = {k: v.clone().requires_grad_(True) for k, v in critic["p"].items()}
    o_critic = (pc, oa.Adam(pc, cfg["critic_lr"], cfg["opti_eps"], cfg["weight_decay"]))
    o_infos, o_cinfo, o_factors, _ = ot.ha_train_hatrpo(o_actors, o_critic, cfg, heads, abufs, cbuf, vn, runner.state_type,
                                                        order, lambda n: np.arange(n))
    for a in range(runner.num_agents):
        np.testing.assert_allclose(runner.actor_buffer[a].factor.cpu().numpy(), o_factors[a], rtol=3e-3, atol=2e-4)
        for k in ("kl", "dist_entropy", "loss_improve", "expected_improve", "ratio"):
            np.testing.assert_allclose(infos[a][k], o_infos[a][k], rtol=5e-3, atol=1e-5, err_msg=f"{k}[{a}]")
        for k, v in runner.actor[a].actor.state_dict().items():
            np.testing.assert_allclose(v.cpu().numpy(), o_actors[a][k].detach().numpy(), rtol=0, atol=3e-4, err_msg=k)
    runner.close()
