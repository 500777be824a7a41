"""Pin the CPU oracle against vectors produced by the unmodified reference (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import algo as oa
from oracle import buffers as ob
from oracle import nets as on
from tests import util as U

torch.set_num_threads(1)


@pytest.mark.parametrize("st", ["EP", "FP"])
def test_insert_masks_bit_exact(st):
    g = U.load(f"insert_{st}")
    T = g["dones"].shape[0]
    for t in range(T):
        masks, active, bad, denv = ob.derive_masks(g["dones"][t], g["bad"][t], st)
        for a in range(3):
            assert np.array_equal(g[f"a{a}.masks"][t + 1], masks[:, a])
            assert np.array_equal(g[f"a{a}.active_masks"][t + 1], active[:, a])
            # rnn states of finished envs are zeroed (on_policy_base_runner.py:358-386)
            assert np.all(g[f"a{a}.rnn_states"][t + 1][denv] == 0)
        if st == "EP":
            assert np.array_equal(g["c.masks"][t + 1], masks[:, 0])
        else:
            assert np.array_equal(g["c.masks"][t + 1], masks)
        assert np.array_equal(g["c.bad_masks"][t + 1], bad)


@pytest.mark.parametrize("name", U.names("gae_"))
def test_gae_bit_exact(name):
    g = U.load(name)
    use_gae, ptl, use_vn = (int(name.split(k)[1][0]) for k in ("_gae", "_ptl", "_vn"))
    vn = None
    if use_vn:
        vn = ob.ValueNormState()
        vn.running_mean, vn.running_mean_sq, vn.debiasing_term = (
            np.float32(g["vn_mean"][0]), np.float32(g["vn_mean_sq"][0]), np.float32(g["vn_debias"]))
    ret, vp = ob.compute_returns(g["in.rewards"], g["in.value_preds"], g["in.masks"], g["in.bad_masks"],
                                 g["next_value"], float(g["gamma"]), float(g["gae_lambda"]),
                                 bool(use_gae), bool(ptl), vn)
    assert np.array_equal(ret[:-1], g["returns"][:-1])
    if use_gae:
        assert np.array_equal(vp, g["value_preds"])
    else:
        assert np.array_equal(ret[-1], g["returns"][-1])
    adv = ob.advantages(g["returns"], g["value_preds"], vn)
    assert np.array_equal(adv, g["advantages"])


def test_valuenorm():
    g = U.load("valuenorm")
    vn = ob.ValueNormState()
    for x, st in zip(g["xs"], g["states"]):
        vn.update(x)
        np.testing.assert_allclose([vn.running_mean, vn.running_mean_sq, vn.debiasing_term], st, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(vn.normalize(g["q"]), g["norm"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(vn.denormalize(g["q"]), g["denorm"], rtol=1e-6, atol=1e-7)


POLICY_CFG = {
    "mlp_disc": ({}, "Discrete"),
    "mlp_box": (dict(hidden_sizes=[32, 32, 32]), "Box"),
    "gru_disc": (dict(use_recurrent_policy=True, hidden_sizes=[16, 16]), "Discrete"),
    "gru_box": (dict(use_recurrent_policy=True, hidden_sizes=[16], recurrent_n=2), "Box"),
    "mlp_disc_tanh": (dict(activation_func="tanh", use_feature_normalization=False), "Discrete"),
}


@pytest.mark.parametrize("tag", sorted(POLICY_CFG))
def test_policy_forward(tag):
    over, head = POLICY_CFG[tag]
    cfg = U.base_args(**over)
    g = U.load(f"policy_{tag}")
    pa, pc = U.params_of(g, "actor/"), U.params_of(g, "critic/")
    t = torch.from_numpy
    avail = t(g["avail"]) if "avail" in g else None
    for mode in ("seq", "row"):
        hx = t(g[f"{mode}.hxs"])
        lp, ent, dist, _ = on.actor_evaluate(pa, cfg, head, t(g["obs"]), hx, t(g["actions"]), t(g["masks"]),
                                             avail, t(g["active"]))
        np.testing.assert_allclose(lp.numpy(), g[f"{mode}.logp"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ent.numpy(), g[f"{mode}.entropy"], rtol=1e-5, atol=1e-6)
        if head == "Discrete":
            np.testing.assert_allclose(dist[1].numpy(), g[f"{mode}.logits"], rtol=1e-5, atol=1e-5)
        else:
            np.testing.assert_allclose(dist[1].numpy(), g[f"{mode}.mean"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(dist[2].numpy(), g[f"{mode}.std"], rtol=1e-6)
        v, hc = on.critic_values(pc, cfg, t(g["cobs"]), hx, t(g["masks"]))
        np.testing.assert_allclose(v.numpy(), g[f"{mode}.values"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(hc.numpy(), g[f"{mode}.hxs_critic_out"], rtol=1e-5, atol=1e-6)
    a, lp, h = on.actor_mode(pa, cfg, head, t(g["obs"]), t(g["row.hxs"]), t(g["masks"]), avail)
    np.testing.assert_allclose(a.numpy(), g["det_action"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lp.numpy(), g["det_logp"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(h.numpy(), g["det_hxs"], rtol=1e-5, atol=1e-6)


def _buf(g, prefix):
    b = U.sub(g, prefix)
    b.setdefault("available_actions", None)
    return b


@pytest.mark.parametrize("name", U.names("single_update_"))
def test_single_update_grads(name):
    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    T, N = cfg["episode_length"], cfg["n_rollout_threads"]
    buf = _buf(g, "a0.")
    p = U.params_of(g, "actor0/", grad=True)
    opt = oa.Adam(p, cfg["lr"], cfg["opti_eps"], cfg["weight_decay"])
    ident = lambda n: np.arange(n)
    batch = next(oa.actor_minibatches(buf, g["adv"], g["factor"], cfg, ident))
    u = oa.happo_update(p, opt, cfg, m["head"], batch)
    np.testing.assert_allclose([u["policy_loss"], u["dist_entropy"], u["actor_grad_norm"], u["ratio"]],
                               g["actor_scalars"], rtol=2e-5, atol=1e-6)
    coef = min(1.0, cfg["max_grad_norm"] / (u["actor_grad_norm"] + 1e-6))
    for k in p:
        np.testing.assert_allclose(u["raw_grads"][k].numpy() * coef, g["grad.actor0/" + k], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(p[k].detach().numpy(), g["out.actor0/" + k], rtol=1e-5, atol=2e-6)
    # critic
    pc = U.params_of(g, "critic/", grad=True)
    copt = oa.Adam(pc, cfg["critic_lr"], cfg["opti_eps"], cfg["weight_decay"])
    vn = ob.ValueNormState()
    vn.running_mean, vn.running_mean_sq, vn.debiasing_term = (np.float32(x) for x in g["vn_in"])
    cb = U.sub(g, "c.")
    cbatch = next(oa.critic_minibatches(cb, cfg, ident))
    cu = oa.critic_update(pc, copt, cfg, cbatch, vn)
    np.testing.assert_allclose([cu["value_loss"], cu["critic_grad_norm"]], g["critic_scalars"], rtol=2e-5)
    coef = min(1.0, cfg["max_grad_norm"] / (cu["critic_grad_norm"] + 1e-6))
    for k in pc:
        np.testing.assert_allclose(cu["raw_grads"][k].numpy() * coef, g["grad.critic/" + k], rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(pc[k].detach().numpy(), g["out.critic/" + k], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", U.names("ha_train_"))
def test_ha_train_iteration(name):
    """Full sequential-agent update: weights, factors, infos, ValueNorm state."""
    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    A = m["A"]
    actors, abufs = [], []
    for a in range(A):
        p = U.params_of(g, f"actor{a}/", grad=True)
        actors.append((p, oa.Adam(p, cfg["lr"], cfg["opti_eps"], cfg["weight_decay"])))
        abufs.append(_buf(g, f"a{a}."))
    pc = U.params_of(g, "critic/", grad=True)
    critic = (pc, oa.Adam(pc, cfg["critic_lr"], cfg["opti_eps"], cfg["weight_decay"]))
    vn = ob.ValueNormState()
    vn.running_mean, vn.running_mean_sq, vn.debiasing_term = (np.float32(x) for x in g["vn_in"])
    cb = U.sub(g, "c.")
    infos, cinfo, factors, _ = oa.ha_train(actors, critic, cfg, [m["head"]] * A, abufs, cb, vn,
                                           m["state_type"], list(range(A)), U.perm_replayer(g))
    for a in range(A):
        np.testing.assert_allclose(factors[a], g[f"out.factor{a}"], rtol=2e-4, atol=1e-5)
        ref = g[f"out.info{a}"]
        got = [infos[a][k] for k in ("policy_loss", "dist_entropy", "actor_grad_norm", "ratio")]
        np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5)
        for k, v in actors[a][0].items():
            np.testing.assert_allclose(v.detach().numpy(), g[f"out.actor{a}/" + k], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose([cinfo["value_loss"], cinfo["critic_grad_norm"]], g["out.cinfo"], rtol=2e-4)
    for k, v in pc.items():
        np.testing.assert_allclose(v.detach().numpy(), g["out.critic/" + k], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose([vn.running_mean, vn.running_mean_sq, vn.debiasing_term], g["out.vn"], rtol=1e-5)


@pytest.mark.parametrize("name", U.names("ma_train_"))
def test_ma_train_iteration(name):
    """OnPolicyMARunner.train of the unmodified reference (MAPPO, with and without parameter sharing) vs oracle.ma_train."""
    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    A, share = m["A"], bool(int(g["share"][0]))
    actors, abufs = [], []
    for a in range(A):
        if share and a > 0:
            actors.append(actors[0])
        else:
            p = U.params_of(g, f"actor{a}/", grad=True)
            actors.append((p, oa.Adam(p, cfg["lr"], cfg["opti_eps"], cfg["weight_decay"])))
        abufs.append(_buf(g, f"a{a}."))
    pc = U.params_of(g, "critic/", grad=True)
    critic = (pc, oa.Adam(pc, cfg["critic_lr"], cfg["opti_eps"], cfg["weight_decay"]))
    vn = ob.ValueNormState()
    vn.running_mean, vn.running_mean_sq, vn.debiasing_term = (np.float32(x) for x in g["vn_in"])
    cb = U.sub(g, "c.")
    infos, cinfo = oa.ma_train(actors, critic, cfg, [m["head"]] * A, abufs, cb, vn, m["state_type"], share, U.perm_replayer(g))
    for a in range(A):
        got = [infos[a][k] for k in ("policy_loss", "dist_entropy", "actor_grad_norm", "ratio")]
        np.testing.assert_allclose(got, g[f"out.info{a}"], rtol=2e-4, atol=2e-5)
        for k, v in actors[a][0].items():
            np.testing.assert_allclose(v.detach().numpy(), g[f"out.actor{a}/" + k], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose([cinfo["value_loss"], cinfo["critic_grad_norm"]], g["out.cinfo"], rtol=2e-4)
    for k, v in pc.items():
        np.testing.assert_allclose(v.detach().numpy(), g["out.critic/" + k], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose([vn.running_mean, vn.running_mean_sq, vn.debiasing_term], g["out.vn"], rtol=1e-5)


@pytest.mark.parametrize("name", ["ha_train_mlp_disc_EP", "ha_train_mlp_box_EP"])
def test_pre_update_logp_sweep_equals_the_first_epochs_forward(name):
    """The identity the device path relies on (hb_ppo_actor_grad_logp, on_policy_ha_runner.py in harl_b200): the
    reference's pre-update evaluate sweep over the whole buffer (on_policy_ha_runner.py:66-83) and the forward of the FIRST
    PPO epoch (happo.py:41-49 on the single whole-buffer minibatch) run the same network with the same weights on the
    same rows -- a row's log-prob does not depend on the order of the rows, so the sweep's numbers are the epoch's, up to
    the permutation.  Shown on the reference's own goldens with the oracle."""
    g = U.load(name)
    cfg, m = U.cfg_of(g), U.meta_of(g)
    assert cfg["actor_num_mini_batch"] == 1
    p = U.params_of(g, "actor0/", grad=True)
    buf = _buf(g, "a0.")
    T, N = buf["actions"].shape[:2]
    sweep = oa.logp_sweep(p, cfg, m["head"], buf)                       # [T*N, ad], time-major rows
    perm = np.random.default_rng(0).permutation(T * N)                  # what the feed-forward generator draws
    adv = np.zeros((T, N, 1), np.float32)
    batch = next(oa.actor_minibatches(buf, adv, np.ones((T, N, 1), np.float32), cfg, lambda n: perm))
    with torch.no_grad():
        lp, _, _, _ = on.actor_evaluate(p, cfg, m["head"], batch["obs"], batch["rnn"], batch["actions"], batch["masks"],
                                        batch.get("avail"), batch["active"])
    np.testing.assert_allclose(lp.numpy(), sweep.numpy()[perm], rtol=0, atol=2e-6)
