"""Generate golden vectors by running the UNMODIFIED reference (PKU-MARL/HARL) on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's own functions are imported from /root/reference and executed on seeded
synthetic inputs; inputs and outputs are written to tests/golden/*.npz (a few hundred KB in
total).  Nothing from the reference is copied into the repo -- only tensors it computed.
The GPU box has no /root/reference: tests read the committed .npz files only.

Shims (the same two SURVEY.md section 8(c) lists): a stub ``tensorboardX`` module (imported
by harl/utils/configs_tools.py:86) and fake ``Box`` / ``Discrete`` space classes (the
reference dispatches on ``__class__.__name__``).
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HARL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
tbx = types.ModuleType("tensorboardX")
tbx.SummaryWriter = object
sys.modules.setdefault("tensorboardX", tbx)

from harl.algorithms.actors.happo import HAPPO  # noqa: E402
from harl.algorithms.critics.v_critic import VCritic  # noqa: E402
from harl.common.buffers.on_policy_actor_buffer import OnPolicyActorBuffer  # noqa: E402
from harl.common.buffers.on_policy_critic_buffer_ep import OnPolicyCriticBufferEP  # noqa: E402
from harl.common.buffers.on_policy_critic_buffer_fp import OnPolicyCriticBufferFP  # noqa: E402
from harl.common.valuenorm import ValueNorm  # noqa: E402
from harl.models.policy_models.stochastic_policy import StochasticPolicy  # noqa: E402
from harl.models.value_function_models.v_net import VNet  # noqa: E402
from harl.runners.on_policy_base_runner import OnPolicyBaseRunner  # noqa: E402
from harl.runners.on_policy_ha_runner import OnPolicyHARunner  # noqa: E402


class Box:
    def __init__(self, n):
        self.shape = (n,)


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


VERSIONS = np.array([torch.__version__, np.__version__])


def base_args(**over):
    a = dict(
        hidden_sizes=[32, 32], activation_func="relu", use_feature_normalization=True,
        initialization_method="orthogonal_", gain=0.01, use_naive_recurrent_policy=False,
        use_recurrent_policy=False, recurrent_n=1, data_chunk_length=4, lr=5e-4, critic_lr=5e-4,
        opti_eps=1e-5, weight_decay=0, std_x_coef=1, std_y_coef=0.5,
        ppo_epoch=3, critic_epoch=3, use_clipped_value_loss=True, clip_param=0.2,
        actor_num_mini_batch=1, critic_num_mini_batch=1, entropy_coef=0.01, value_loss_coef=1,
        use_max_grad_norm=True, max_grad_norm=10.0, use_gae=True, gamma=0.99, gae_lambda=0.95,
        use_huber_loss=True, use_policy_active_masks=True, huber_delta=10.0,
        action_aggregation="prod", share_param=False, fixed_order=True,
        episode_length=8, n_rollout_threads=6, use_valuenorm=True, use_proper_time_limits=True,
        kl_threshold=0.01, ls_step=10, accept_ratio=0.5, backtrack_coeff=0.8,
    )
    a.update(over)
    return a


def sd_np(module, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def perturb(module, g, scale=0.1):
    """Move every parameter off its init so LN affines / biases / log_std matter."""
    with torch.no_grad():
        for p in module.parameters():
            p.add_(scale * torch.randn(p.shape, generator=g))


def save(name, **arrs):
    arrs["versions"] = VERSIONS
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print("wrote", name, sum(np.asarray(v).nbytes for v in arrs.values()), "bytes")


# ------------------------------------------------------------------ random rollout data
def fill_buffers(rng, args, A, od, sd, act_space, state_type, with_deaths=True):
    """Drive the reference runner's own insert() with random env outputs; return buffers."""
    T, N = args["episode_length"], args["n_rollout_threads"]
    h, R = args["hidden_sizes"][-1], args["recurrent_n"]
    disc = act_space.__class__.__name__ == "Discrete"
    abufs = [OnPolicyActorBuffer(args, Box(od), act_space) for _ in range(A)]
    cbuf = (OnPolicyCriticBufferEP(args, Box(sd)) if state_type == "EP"
            else OnPolicyCriticBufferFP(args, Box(sd), A))
    fake = SimpleNamespace(num_agents=A, recurrent_n=R, rnn_hidden_size=h, state_type=state_type,
                           algo_args={"train": {"n_rollout_threads": N}},
                           actor_buffer=abufs, critic_buffer=cbuf)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)

    def avail_sample():
        av = (rng.random((N, A, act_space.n)) < 0.7).astype(np.float32)
        av[..., 1] = 1.0
        return av

    for a in range(A):
        abufs[a].obs[0] = f(N, od)
        if disc:
            abufs[a].available_actions[0] = avail_sample()[:, a]
    cbuf.share_obs[0] = f(N, sd) if state_type == "EP" else f(N, A, sd)
    dead = np.zeros((N, A), bool)
    log = dict(dones=[], bad=[])
    for t in range(T):
        if disc:
            actions = rng.integers(0, act_space.n, (N, A, 1)).astype(np.float32)
            ad = 1
        else:
            ad = act_space.shape[0]
            actions = f(N, A, ad)
        logp = -np.abs(f(N, A, ad)) - 0.5
        values = f(N, 1) if state_type == "EP" else f(N, A, 1)
        rnn = f(N, A, R, h)
        rnn_c = f(N, R, h) if state_type == "EP" else f(N, A, R, h)
        rew = np.repeat(f(N, 1, 1), A, axis=1)
        if with_deaths:
            dead |= rng.random((N, A)) < 0.12
        env_done = dead.all(1) | (rng.random(N) < 0.1)
        dones = dead | env_done[:, None]
        trunc = env_done & (rng.random(N) < 0.5)
        infos = [[({"bad_transition": True} if trunc[n] else {}) for _ in range(A)] for n in range(N)]
        data = (f(N, A, od), f(N, A, sd), rew, dones.copy(), infos,
                avail_sample() if disc else np.array([None] * N), values, actions, logp, rnn, rnn_c)
        OnPolicyBaseRunner.insert(fake, data)
        log["dones"].append(dones.copy())
        log["bad"].append(np.repeat(trunc[:, None], A, 1))
        dead[env_done] = False
    return abufs, cbuf, log


def abuf_np(b, prefix):
    d = {prefix + k: getattr(b, k).copy() for k in
         ("obs", "rnn_states", "actions", "action_log_probs", "masks", "active_masks")}
    if b.available_actions is not None:
        d[prefix + "available_actions"] = b.available_actions.copy()
    return d


def cbuf_np(b, prefix="c."):
    return {prefix + k: getattr(b, k).copy() for k in
            ("share_obs", "rnn_states_critic", "value_preds", "returns", "rewards", "masks", "bad_masks")}


# ------------------------------------------------------------------ cases
def case_insert():
    for st in ("EP", "FP"):
        rng = np.random.default_rng(11)
        args = base_args(hidden_sizes=[8])
        ab, cb, log = fill_buffers(rng, args, 3, 5, 7, Discrete(4), st)
        out = dict(dones=np.array(log["dones"]), bad=np.array(log["bad"]))
        for a in range(3):
            out.update({f"a{a}.masks": ab[a].masks, f"a{a}.active_masks": ab[a].active_masks,
                        f"a{a}.rnn_states": ab[a].rnn_states})
        out.update({"c.masks": cb.masks, "c.bad_masks": cb.bad_masks, "c.rnn_states_critic": cb.rnn_states_critic})
        save(f"insert_{st}", **out)


def case_gae():
    for st in ("EP", "FP"):
        for use_gae in (True, False):
            for ptl in (True, False):
                for use_vn in (True, False):
                    rng = np.random.default_rng(5)
                    args = base_args(use_gae=use_gae, use_proper_time_limits=ptl, hidden_sizes=[8],
                                     episode_length=16, n_rollout_threads=5)
                    ab, cb, _ = fill_buffers(rng, args, 3, 4, 6, Box(2), st)
                    vn = None
                    if use_vn:
                        vn = ValueNorm(1)
                        vn.update(rng.standard_normal((64, 1)).astype(np.float32) * 3 + 1)
                        vn.update(rng.standard_normal((64, 1)).astype(np.float32) * 2 - 1)
                    nv = rng.standard_normal(cb.value_preds[-1].shape).astype(np.float32)
                    inp = cbuf_np(cb, "in.")
                    cb.compute_returns(nv, vn)
                    out = dict(inp, next_value=nv, returns=cb.returns, value_preds=cb.value_preds,
                               gamma=args["gamma"], gae_lambda=args["gae_lambda"])
                    if use_vn:
                        out.update(vn_mean=vn.running_mean.numpy(), vn_mean_sq=vn.running_mean_sq.numpy(),
                                   vn_debias=vn.debiasing_term.numpy(),
                                   denorm_vp=vn.denormalize(cb.value_preds[:-1]))
                        adv = cb.returns[:-1] - vn.denormalize(cb.value_preds[:-1])
                    else:
                        adv = cb.returns[:-1] - cb.value_preds[:-1]
                    out["advantages"] = adv
                    save(f"gae_{st}_gae{int(use_gae)}_ptl{int(ptl)}_vn{int(use_vn)}", **out)


def case_valuenorm():
    rng = np.random.default_rng(2)
    vn = ValueNorm(1)
    xs = [rng.standard_normal((50, 1)).astype(np.float32) * s + m for s, m in ((3, 1), (0.01, 0), (10, -4))]
    st = []
    q = rng.standard_normal((9, 1)).astype(np.float32)
    for x in xs:
        vn.update(x)
        st.append([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()])
    save("valuenorm", xs=np.array(xs), states=np.array(st, np.float32), q=q,
         norm=vn.normalize(q).numpy(), denorm=vn.denormalize(q))


def case_policy():
    for tag, over, act_space in (
        ("mlp_disc", {}, Discrete(5)),
        ("mlp_box", dict(hidden_sizes=[32, 32, 32]), Box(3)),
        ("gru_disc", dict(use_recurrent_policy=True, hidden_sizes=[16, 16]), Discrete(6)),
        ("gru_box", dict(use_recurrent_policy=True, hidden_sizes=[16], recurrent_n=2), Box(2)),
        ("mlp_disc_tanh", dict(activation_func="tanh", use_feature_normalization=False), Discrete(4)),
    ):
        torch.manual_seed(3)
        g = torch.Generator().manual_seed(4)
        rng = np.random.default_rng(6)
        args = base_args(**over)
        od, T, N = 7, 5, 4
        h, R = args["hidden_sizes"][-1], args["recurrent_n"]
        pol = StochasticPolicy(args, Box(od), act_space)
        perturb(pol, g)
        vnet = VNet(args, Box(od + 2))
        perturb(vnet, g)
        disc = act_space.__class__.__name__ == "Discrete"
        f = lambda *s: rng.standard_normal(s).astype(np.float32)
        B = T * N
        obs, cobs = f(B, od), f(B, od + 2)
        masks = (rng.random((B, 1)) > 0.2).astype(np.float32)
        active = (rng.random((B, 1)) > 0.2).astype(np.float32)
        if disc:
            avail = (rng.random((B, act_space.n)) < 0.7).astype(np.float32)
            avail[:, 1] = 1
            acts = np.array([[rng.choice(np.flatnonzero(avail[i]))] for i in range(B)], np.float32)
        else:
            avail, acts = None, f(B, act_space.shape[0])
        out = dict(obs=obs, cobs=cobs, masks=masks, active=active, actions=acts)
        if disc:
            out["avail"] = avail
        # sequence mode (rows = T*N time-major, hxs [N,R,h]) and row mode (hxs [B,R,h])
        for mode, hx in (("seq", f(N, R, h)), ("row", f(B, R, h))):
            lp, ent, dist = pol.evaluate_actions(obs, hx, acts, masks, avail, active)
            out[f"{mode}.hxs"] = hx
            out[f"{mode}.logp"] = lp.detach().numpy()
            out[f"{mode}.entropy"] = ent.detach().numpy()
            if disc:
                out[f"{mode}.logits"] = dist.logits.detach().numpy()
            else:
                out[f"{mode}.mean"] = dist.loc.detach().numpy()
                out[f"{mode}.std"] = dist.scale.detach().numpy()
            v, hc = vnet(cobs, hx, masks)
            out[f"{mode}.values"] = v.detach().numpy()
            out[f"{mode}.hxs_critic_out"] = hc.detach().numpy()
        a_det, lp_det, h_det = pol(obs, out["row.hxs"], masks, avail, deterministic=True)
        out.update(det_action=a_det.detach().numpy().astype(np.float32), det_logp=lp_det.detach().numpy(),
                   det_hxs=h_det.detach().numpy())
        out.update(sd_np(pol, "actor/"))
        out.update(sd_np(vnet, "critic/"))
        save(f"policy_{tag}", **out)


class PermRecorder:
    """Record (and later replay) every torch.randperm the reference draws."""

    def __init__(self):
        self.log = []
        self._orig = torch.randperm

    def __enter__(self):
        def rec(n, *a, **k):
            p = self._orig(n, *a, **k)
            self.log.append(p.numpy().copy())
            return p
        torch.randperm = rec
        return self

    def __exit__(self, *a):
        torch.randperm = self._orig


def case_ha_train():
    """Full OnPolicyHARunner.train() through the unmodified reference classes."""
    for tag, over, act_space, st, A in (
        ("mlp_disc_EP", {}, Discrete(5), "EP", 3),
        ("mlp_box_EP", dict(hidden_sizes=[32, 32, 32], clip_param=0.05), Box(2), "EP", 2),
        ("mlp_disc_mb2_EP", dict(actor_num_mini_batch=2, critic_num_mini_batch=2), Discrete(5), "EP", 2),
        ("gru_disc_FP", dict(use_recurrent_policy=True, hidden_sizes=[16, 16], gamma=0.95), Discrete(6), "FP", 3),
        ("gru_box_EP", dict(use_recurrent_policy=True, hidden_sizes=[16]), Box(2), "EP", 2),
        ("naive_gru_disc_EP", dict(use_naive_recurrent_policy=True, hidden_sizes=[16]), Discrete(4), "EP", 2),
    ):
        torch.manual_seed(7)
        g = torch.Generator().manual_seed(8)
        rng = np.random.default_rng(9)
        args = base_args(**over)
        od, sd = 6, 9
        ab, cb, _ = fill_buffers(rng, args, A, od, sd, act_space, st)
        actors = [HAPPO(args, Box(od), act_space) for _ in range(A)]
        critic = VCritic(args, Box(sd))
        for a in actors:
            perturb(a.actor, g)
        perturb(critic.critic, g)
        vn = ValueNorm(1)
        vn.update(rng.standard_normal((64, 1)).astype(np.float32) * 2 + 0.5)
        nv = rng.standard_normal(cb.value_preds[-1].shape).astype(np.float32)
        cb.compute_returns(nv, vn)
        out = {}
        for a in range(A):
            out.update(abuf_np(ab[a], f"a{a}."))
            out.update(sd_np(actors[a].actor, f"actor{a}/"))
        out.update(cbuf_np(cb))
        out.update(sd_np(critic.critic, "critic/"))
        out["vn_in"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()], np.float32)
        fake = SimpleNamespace(algo_args={"train": args}, value_normalizer=vn, critic_buffer=cb, actor_buffer=ab,
                               actor=actors, critic=critic, state_type=st, num_agents=A, fixed_order=True,
                               action_aggregation=args["action_aggregation"])
        for x in actors:
            x.prep_training()
        critic.prep_training()
        with PermRecorder() as pr:
            ainfos, cinfo = OnPolicyHARunner.train(fake)
        out["n_perms"] = len(pr.log)
        for i, p in enumerate(pr.log):
            out[f"perm{i}"] = p
        for a in range(A):
            out.update(sd_np(actors[a].actor, f"out.actor{a}/"))
            out[f"out.factor{a}"] = ab[a].factor
            out[f"out.info{a}"] = np.array([float(ainfos[a][k]) for k in
                                            ("policy_loss", "dist_entropy", "actor_grad_norm", "ratio")], np.float64)
        out.update(sd_np(critic.critic, "out.critic/"))
        out["out.cinfo"] = np.array([float(cinfo["value_loss"]), float(cinfo["critic_grad_norm"])], np.float64)
        out["out.vn"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()], np.float32)
        out["cfg_keys"] = np.array(sorted(over.keys()))
        out["cfg_vals"] = np.array([repr(over[k]) for k in sorted(over.keys())])
        out["meta"] = np.array([tag, act_space.__class__.__name__, st, str(A), str(od), str(sd),
                                str(act_space.n if hasattr(act_space, "n") else act_space.shape[0])])
        save(f"ha_train_{tag}", **out)


def case_ma_train():
    """Full OnPolicyMARunner.train() (MAPPO, with and without parameter sharing) through the unmodified reference."""
    from harl.algorithms.actors.mappo import MAPPO
    from harl.runners.on_policy_ma_runner import OnPolicyMARunner

    for tag, over, act_space, st, A, share in (
        ("disc_EP", {}, Discrete(5), "EP", 3, False),
        ("disc_EP_share", {}, Discrete(5), "EP", 3, True),
        ("box_FP_share_mb2", dict(hidden_sizes=[32, 32, 32], actor_num_mini_batch=2, critic_num_mini_batch=2), Box(2), "FP", 2, True),
    ):
        torch.manual_seed(17)
        g = torch.Generator().manual_seed(18)
        rng = np.random.default_rng(19)
        args = base_args(**over)
        od, sd = 6, 9
        ab, cb, _ = fill_buffers(rng, args, A, od, sd, act_space, st)
        for b in ab:
            b.factor = None
        if share:
            first = MAPPO(args, Box(od), act_space)
            actors = [first] * A
        else:
            actors = [MAPPO(args, Box(od), act_space) for _ in range(A)]
        critic = VCritic(args, Box(sd))
        for a in (actors[:1] if share else actors):
            perturb(a.actor, g)
        perturb(critic.critic, g)
        vn = ValueNorm(1)
        vn.update(rng.standard_normal((64, 1)).astype(np.float32) * 2 + 0.5)
        nv = rng.standard_normal(cb.value_preds[-1].shape).astype(np.float32)
        cb.compute_returns(nv, vn)
        out = {}
        for a in range(A):
            out.update(abuf_np(ab[a], f"a{a}."))
            out.update(sd_np(actors[a].actor, f"actor{a}/"))
        out.update(cbuf_np(cb))
        out.update(sd_np(critic.critic, "critic/"))
        out["vn_in"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()], np.float32)
        fake = SimpleNamespace(algo_args={"train": args}, value_normalizer=vn, critic_buffer=cb, actor_buffer=ab,
                               actor=actors, critic=critic, state_type=st, num_agents=A, share_param=share)
        for x in actors:
            x.prep_training()
        critic.prep_training()
        with PermRecorder() as pr:
            ainfos, cinfo = OnPolicyMARunner.train(fake)
        # the runner draws one randperm(num_agents) that only orders the info list: drop it from the replay log
        perms = [q for q in pr.log if not (share and len(q) == A)]
        out["n_perms"] = len(perms)
        for i, q in enumerate(perms):
            out[f"perm{i}"] = q
        for a in range(A):
            out.update(sd_np(actors[a].actor, f"out.actor{a}/"))
            out[f"out.info{a}"] = np.array([float(ainfos[a][k]) for k in
                                            ("policy_loss", "dist_entropy", "actor_grad_norm", "ratio")], np.float64)
        out.update(sd_np(critic.critic, "out.critic/"))
        out["out.cinfo"] = np.array([float(cinfo["value_loss"]), float(cinfo["critic_grad_norm"])], np.float64)
        out["out.vn"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()], np.float32)
        out["share"] = np.array([int(share)])
        out["cfg_keys"] = np.array(sorted(over.keys()))
        out["cfg_vals"] = np.array([repr(over[k]) for k in sorted(over.keys())])
        out["meta"] = np.array([tag, act_space.__class__.__name__, st, str(A), str(od), str(sd),
                                str(act_space.n if hasattr(act_space, "n") else act_space.shape[0])])
        save(f"ma_train_{tag}", **out)


def case_checkpoint():
    """A checkpoint directory written by the unmodified reference's OnPolicyBaseRunner.save() (on_policy_base_runner.py:
    724-740) and what the reference computes from it: deterministic actions of every actor and the critic's values on a
    fixed batch.  tests/golden/ckpt_ref/*.pt are the reference's own torch.save files."""
    torch.manual_seed(41)
    g = torch.Generator().manual_seed(42)
    rng = np.random.default_rng(43)
    args = base_args()
    A, od, sd, na, n = 3, 7, 9, 5, 12
    actors = [HAPPO(args, Box(od), Discrete(na)) for _ in range(A)]
    critic = VCritic(args, Box(sd))
    for a in actors:
        perturb(a.actor, g, 0.3)
    perturb(critic.critic, g, 0.3)
    vn = ValueNorm(1)
    vn.update(rng.standard_normal((64, 1)).astype(np.float32) * 2 + 0.5)
    d = os.path.join(HERE, "ckpt_ref")
    os.makedirs(d, exist_ok=True)
    fake = SimpleNamespace(num_agents=A, actor=actors, critic=critic, value_normalizer=vn, save_dir=d)
    OnPolicyBaseRunner.save(fake)
    obs = rng.standard_normal((n, A, od)).astype(np.float32)
    share = rng.standard_normal((n, sd)).astype(np.float32)
    avail = (rng.random((n, A, na)) < 0.6).astype(np.float32)
    avail[..., 2] = 1.0
    rnn = np.zeros((n, 1, 32), np.float32)
    masks = np.ones((n, 1), np.float32)
    out = dict(obs=obs, share_obs=share, avail=avail)
    for a in range(A):
        for x in (actors[a], critic):
            x.prep_rollout()
        act, _ = actors[a].act(obs[:, a], rnn, masks, avail[:, a], deterministic=True)
        out[f"det_action{a}"] = act.detach().numpy()
        _, lp, _ = actors[a].get_actions(obs[:, a], rnn, masks, avail[:, a], deterministic=True)
        out[f"det_logp{a}"] = lp.detach().numpy()
    v, _ = critic.get_values(share, rnn, masks)
    out["values"] = v.detach().numpy()
    out["values_denorm"] = vn.denormalize(v.detach()) if not isinstance(vn.denormalize(v.detach()), torch.Tensor) else vn.denormalize(v.detach()).numpy()
    out["vn"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()], np.float32)
    save("checkpoint_ref", **out)
    print("files:", sorted(os.listdir(d)))


ACTIVATIONS = ("sigmoid", "tanh", "relu", "leaky_relu", "selu", "hardswish", "identity")   # models_tools.py:28-50


def case_single_update(cases=None):
    """One HAPPO.update and one VCritic.update: raw gradients before clipping."""
    cases = cases or (("disc", {}, Discrete(5)), ("box", dict(hidden_sizes=[32, 32, 32]), Box(3)))
    for tag, over, act_space in cases:
        torch.manual_seed(21)
        g = torch.Generator().manual_seed(22)
        rng = np.random.default_rng(23)
        args = base_args(ppo_epoch=1, critic_epoch=1, **over)
        od, sd, A = 6, 9, 1
        ab, cb, _ = fill_buffers(rng, args, A, od, sd, act_space, "EP")
        actor = HAPPO(args, Box(od), act_space)
        critic = VCritic(args, Box(sd))
        perturb(actor.actor, g)
        perturb(critic.critic, g)
        vn = ValueNorm(1)
        vn.update(rng.standard_normal((64, 1)).astype(np.float32))
        cb.compute_returns(rng.standard_normal((args["n_rollout_threads"], 1)).astype(np.float32), vn)
        T, N = args["episode_length"], args["n_rollout_threads"]
        factor = (1 + 0.1 * rng.standard_normal((T, N, 1))).astype(np.float32)
        adv = rng.standard_normal((T, N, 1)).astype(np.float32)
        ab[0].update_factor(factor)
        out = dict(abuf_np(ab[0], "a0."), **cbuf_np(cb), factor=factor, adv=adv)
        out.update(sd_np(actor.actor, "actor0/"))
        out.update(sd_np(critic.critic, "critic/"))
        out["vn_in"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()], np.float32)
        # identity permutation so the batch is the buffer in time-major order
        orig = torch.randperm
        torch.randperm = lambda n, *a, **k: torch.arange(n)
        try:
            sample = next(ab[0].feed_forward_generator_actor(adv, 1))
            pl, ent, gn, imp = actor.update(sample)
            for n_, p in actor.actor.named_parameters():
                out["grad.actor0/" + n_] = p.grad.numpy().copy()  # post-clip grads
            out["actor_scalars"] = np.array([pl.item(), ent.item(), float(gn), imp.mean().item()], np.float64)
            csample = next(cb.feed_forward_generator_critic(1))
            vl, cgn = critic.update(csample, vn)
            for n_, p in critic.critic.named_parameters():
                out["grad.critic/" + n_] = p.grad.numpy().copy()
            out["critic_scalars"] = np.array([vl.item(), float(cgn)], np.float64)
        finally:
            torch.randperm = orig
        out.update(sd_np(actor.actor, "out.actor0/"))
        out.update(sd_np(critic.critic, "out.critic/"))
        out["meta"] = np.array([tag, act_space.__class__.__name__, "EP", "1", str(od), str(sd),
                                str(act_space.n if hasattr(act_space, "n") else act_space.shape[0])])
        out["cfg_keys"] = np.array(sorted(list(over.keys()) + ["ppo_epoch", "critic_epoch"]))
        over2 = dict(over, ppo_epoch=1, critic_epoch=1)
        out["cfg_vals"] = np.array([repr(over2[k]) for k in sorted(over2.keys())])
        save(f"single_update_{tag}", **out)


def _named_flat(actor, flat):
    """Split a flat vector in ``actor.parameters()`` order into {name: array} (trpo_util.py:28-35 order)."""
    out, i = {}, 0
    for n_, p in actor.named_parameters():
        k = p.numel()
        out[n_] = flat[i:i + k].detach().numpy().reshape(p.shape).copy()
        i += k
    return out


def case_hatrpo_parts():
    """The pieces of one HATRPO.update (hatrpo.py:37-194): surrogate gradient, one Fisher-vector product, the
    conjugate-gradient step direction, and the whole update incl. the backtracking line search."""
    from harl.algorithms.actors.hatrpo import HATRPO
    from harl.utils import trpo_util as tu

    for tag, over, act_space in (
        ("disc", {}, Discrete(5)),
        ("box", dict(hidden_sizes=[32, 32, 32], action_aggregation="mean"), Box(3)),
        ("disc_nomask", dict(use_policy_active_masks=False, activation_func="tanh", kl_threshold=0.001), Discrete(4)),
        ("disc_backtrack", dict(accept_ratio=0.75), Discrete(5)),         # two trials rejected, the third accepted
        ("box_reject", dict(accept_ratio=5.0, ls_step=3), Box(2)),        # never accepted: parameters restored
        ("gru_disc", dict(use_recurrent_policy=True, hidden_sizes=[16, 16], data_chunk_length=4), Discrete(6)),
        ("gru2_disc", dict(use_recurrent_policy=True, hidden_sizes=[16], recurrent_n=2, data_chunk_length=8), Discrete(4)),
        # synthetic-SMAC head / trunk widths (BASELINE configs[3]): 12 actions, hidden 64
        ("disc12_h64", dict(hidden_sizes=[64, 64]), Discrete(12)),
        ("gru_disc12_h64", dict(use_recurrent_policy=True, hidden_sizes=[64, 64], data_chunk_length=4), Discrete(12)),
    ):
        torch.manual_seed(31)
        g = torch.Generator().manual_seed(32)
        rng = np.random.default_rng(33)
        args = base_args(episode_length=16, n_rollout_threads=8, **over)
        od, sd = 6, 9
        ab, cb, _ = fill_buffers(rng, args, 1, od, sd, act_space, "EP")
        actor = HATRPO(args, Box(od), act_space)
        perturb(actor.actor, g)
        T, N = args["episode_length"], args["n_rollout_threads"]
        factor = (1 + 0.1 * rng.standard_normal((T, N, 1))).astype(np.float32)
        adv = rng.standard_normal((T, N, 1)).astype(np.float32)
        ab[0].update_factor(factor)
        # old log-probs as the rollout would have stored them: the current policy's own (ratio == 1 at theta_old)
        fl = lambda a: a.reshape(T * N, *a.shape[2:])
        with torch.no_grad():
            lp, _, _ = actor.evaluate_actions(fl(ab[0].obs[:-1]),
                                              ab[0].rnn_states[0] if args["use_recurrent_policy"] else fl(ab[0].rnn_states[:-1]),
                                              fl(ab[0].actions),
                                              fl(ab[0].masks[:-1]),
                                              fl(ab[0].available_actions[:-1]) if ab[0].available_actions is not None else None,
                                              fl(ab[0].active_masks[:-1]))
        noise = 0.05 * rng.standard_normal(lp.shape).astype(np.float32)
        ab[0].action_log_probs[:] = (lp.numpy() + noise).reshape(ab[0].action_log_probs.shape)
        out = dict(abuf_np(ab[0], "a0."), factor=factor, adv=adv)
        out.update(sd_np(actor.actor, "actor0/"))
        orig = torch.randperm
        torch.randperm = lambda n, *a, **k: torch.arange(n)
        try:
            if args["use_recurrent_policy"]:  # identity chunk order: the batch is every chunk, step-major
                sample = next(ab[0].recurrent_generator_actor(adv, 1, args["data_chunk_length"]))
            else:
                sample = next(ab[0].feed_forward_generator_actor(adv, 1))
        finally:
            torch.randperm = orig
        (obs_b, rnn_b, act_b, masks_b, active_b, old_lp_b, adv_b, avail_b, factor_b) = sample
        tp = dict(dtype=torch.float32)
        lp, ent, _ = actor.evaluate_actions(obs_b, rnn_b, act_b, masks_b, avail_b, active_b)
        ratio = getattr(torch, args["action_aggregation"])(torch.exp(lp - torch.from_numpy(old_lp_b)), dim=-1, keepdim=True)
        inner = torch.sum(ratio * torch.from_numpy(factor_b) * torch.from_numpy(adv_b), dim=-1, keepdim=True)
        am = torch.from_numpy(active_b)
        loss = (inner * am).sum() / am.sum() if args["use_policy_active_masks"] else inner.mean()
        lg = tu.flat_grad(torch.autograd.grad(loss, actor.actor.parameters(), allow_unused=True))
        out["loss"] = np.array([loss.item()], np.float64)
        for k, v in _named_flat(actor.actor, lg).items():
            out["loss_grad/" + k] = v
        vec = torch.randn(lg.shape, generator=g)
        fvp = tu.fisher_vector_product(actor.actor, obs_b, rnn_b, act_b, masks_b, avail_b, active_b, vec)
        for k, v in _named_flat(actor.actor, vec).items():
            out["vec/" + k] = v
        for k, v in _named_flat(actor.actor, fvp).items():
            out["fvp/" + k] = v
        sdir = tu.conjugate_gradient(actor.actor, obs_b, rnn_b, act_b, masks_b, avail_b, active_b, lg.data, nsteps=10,
                                     device=torch.device("cpu"))
        for k, v in _named_flat(actor.actor, sdir).items():
            out["step_dir/" + k] = v
        kl, li, ei, de, rt = actor.update(sample)
        out["update_scalars"] = np.array([float(kl.detach()), float(li), float(np.asarray(ei).reshape(-1)[0]),
                                          float(de.detach()), float(rt.detach().mean())], np.float64)
        out.update(sd_np(actor.actor, "out.actor0/"))
        out["meta"] = np.array([tag, act_space.__class__.__name__, "EP", "1", str(od), str(sd),
                                str(act_space.n if hasattr(act_space, "n") else act_space.shape[0])])
        over2 = dict(over, episode_length=16, n_rollout_threads=8)
        out["cfg_keys"] = np.array(sorted(over2.keys()))
        out["cfg_vals"] = np.array([repr(over2[k]) for k in sorted(over2.keys())])
        save(f"hatrpo_parts_{tag}", **out)


def case_hatrpo_train():
    """Full OnPolicyHARunner.train() with HATRPO actors (hatrpo.py:196-247) through the unmodified reference."""
    from harl.algorithms.actors.hatrpo import HATRPO

    for tag, over, act_space, st, A in (
        ("mlp_disc_EP", {}, Discrete(5), "EP", 3),
        ("mlp_box_FP", dict(hidden_sizes=[32, 32, 32]), Box(2), "FP", 2),
        ("gru_disc_FP", dict(use_recurrent_policy=True, hidden_sizes=[16, 16], data_chunk_length=4, gamma=0.95), Discrete(6), "FP", 3),
    ):
        torch.manual_seed(41)
        g = torch.Generator().manual_seed(42)
        rng = np.random.default_rng(43)
        args = base_args(episode_length=16, n_rollout_threads=8, **over)
        od, sd = 6, 9
        ab, cb, _ = fill_buffers(rng, args, A, od, sd, act_space, st)
        actors = [HATRPO(args, Box(od), act_space) for _ in range(A)]
        critic = VCritic(args, Box(sd))
        for a in actors:
            perturb(a.actor, g)
        perturb(critic.critic, g)
        # stored log-probs near the current policies' own, as after a rollout
        T, N = args["episode_length"], args["n_rollout_threads"]
        fl = lambda x: x.reshape(T * N, *x.shape[2:])
        for a in range(A):
            with torch.no_grad():
                lp, _, _ = actors[a].evaluate_actions(
                    fl(ab[a].obs[:-1]), ab[a].rnn_states[0] if args["use_recurrent_policy"] else fl(ab[a].rnn_states[:-1]),
                    fl(ab[a].actions), fl(ab[a].masks[:-1]),
                    fl(ab[a].available_actions[:-1]) if ab[a].available_actions is not None else None,
                    fl(ab[a].active_masks[:-1]))
            ab[a].action_log_probs[:] = (lp.numpy() + 0.05 * rng.standard_normal(lp.shape).astype(np.float32)).reshape(
                ab[a].action_log_probs.shape)
        vn = ValueNorm(1)
        vn.update(rng.standard_normal((64, 1)).astype(np.float32) * 2 + 0.5)
        nv = rng.standard_normal(cb.value_preds[-1].shape).astype(np.float32)
        cb.compute_returns(nv, vn)
        out = {}
        for a in range(A):
            out.update(abuf_np(ab[a], f"a{a}."))
            out.update(sd_np(actors[a].actor, f"actor{a}/"))
        out.update(cbuf_np(cb))
        out.update(sd_np(critic.critic, "critic/"))
        out["vn_in"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()], np.float32)
        fake = SimpleNamespace(algo_args={"train": args}, value_normalizer=vn, critic_buffer=cb, actor_buffer=ab,
                               actor=actors, critic=critic, state_type=st, num_agents=A, fixed_order=True,
                               action_aggregation=args["action_aggregation"])
        with PermRecorder() as pr:
            ainfos, cinfo = OnPolicyHARunner.train(fake)
        out["n_perms"] = len(pr.log)
        for i, p in enumerate(pr.log):
            out[f"perm{i}"] = p
        for a in range(A):
            out.update(sd_np(actors[a].actor, f"out.actor{a}/"))
            out[f"out.factor{a}"] = ab[a].factor
            tof = lambda v: float(v.detach().reshape(-1)[0]) if torch.is_tensor(v) else float(np.asarray(v).reshape(-1)[0])
            out[f"out.info{a}"] = np.array([tof(ainfos[a][k]) for k in
                                            ("kl", "dist_entropy", "loss_improve", "expected_improve", "ratio")], np.float64)
        out.update(sd_np(critic.critic, "out.critic/"))
        out["out.cinfo"] = np.array([float(cinfo["value_loss"]), float(cinfo["critic_grad_norm"])], np.float64)
        out["out.vn"] = np.array([vn.running_mean.item(), vn.running_mean_sq.item(), vn.debiasing_term.item()], np.float32)
        over2 = dict(over, episode_length=16, n_rollout_threads=8)
        out["cfg_keys"] = np.array(sorted(over2.keys()))
        out["cfg_vals"] = np.array([repr(over2[k]) for k in sorted(over2.keys())])
        out["meta"] = np.array([tag, act_space.__class__.__name__, st, str(A), str(od), str(sd),
                                str(act_space.n if hasattr(act_space, "n") else act_space.shape[0])])
        save(f"hatrpo_train_{tag}", **out)


def case_generators():
    """Which buffer rows the reference's minibatch generators gather: every buffer element encodes its own (t, n[, a]),
    so the yielded batches ARE the index maps.  Actor buffer + EP and FP critic buffers, the
    three generator kinds, two minibatches each, with the permutation the reference drew."""
    T, N, A, L = 8, 6, 3, 4
    args = base_args(episode_length=T, n_rollout_threads=N, hidden_sizes=[8], data_chunk_length=L,
                     use_recurrent_policy=True)
    out = dict(T=np.array(T), N=np.array(N), A=np.array(A), L=np.array(L))
    enc = (np.arange(T + 1)[:, None] * 1000 + np.arange(N)[None, :]).astype(np.float32)        # t*1000 + n
    ab = OnPolicyActorBuffer(args, Box(3), Discrete(4))
    ab.obs[:] = enc[:, :, None]
    ab.rnn_states[:] = enc[:, :, None, None]
    ab.masks[:] = enc[:, :, None]
    ab.active_masks[:] = enc[:, :, None]
    ab.actions[:] = enc[:T, :, None]
    ab.action_log_probs[:] = enc[:T, :, None]
    adv = enc[:T, :, None].copy()
    gens = dict(ff=lambda: ab.feed_forward_generator_actor(adv, 2), naive=lambda: ab.naive_recurrent_generator_actor(adv, 2),
                chunk=lambda: ab.recurrent_generator_actor(adv, 2, L))
    for kind, mk in gens.items():
        with PermRecorder() as pr:
            batches = list(mk())
        out[f"actor.{kind}.perm"] = pr.log[0]
        for i, b in enumerate(batches):
            out[f"actor.{kind}.{i}.obs"] = b[0][:, 0]
            out[f"actor.{kind}.{i}.rnn"] = b[1][:, 0, 0]
            out[f"actor.{kind}.{i}.actions"] = b[2][:, 0]
            out[f"actor.{kind}.{i}.masks"] = b[3][:, 0]
            out[f"actor.{kind}.{i}.adv"] = b[6][:, 0]
    for st in ("EP", "FP"):
        if st == "EP":
            cb = OnPolicyCriticBufferEP(args, Box(5))
            code = enc
            cb.share_obs[:] = code[:, :, None]
            cb.rnn_states_critic[:] = code[:, :, None, None]
            for k in ("value_preds", "returns", "masks"):
                getattr(cb, k)[:] = code[:, :, None]
        else:
            cb = OnPolicyCriticBufferFP(args, Box(5), A)
            code = (np.arange(T + 1)[:, None, None] * 10000 + np.arange(N)[None, :, None] * 10
                    + np.arange(A)[None, None, :]).astype(np.float32)                            # t*10000 + n*10 + a
            cb.share_obs[:] = code[:, :, :, None]
            cb.rnn_states_critic[:] = code[:, :, :, None, None]
            for k in ("value_preds", "returns", "masks"):
                getattr(cb, k)[:] = code[:, :, :, None]
        gens = dict(ff=lambda: cb.feed_forward_generator_critic(2), naive=lambda: cb.naive_recurrent_generator_critic(2),
                    chunk=lambda: cb.recurrent_generator_critic(2, L))
        for kind, mk in gens.items():
            with PermRecorder() as pr:
                batches = list(mk())
            out[f"critic{st}.{kind}.perm"] = pr.log[0]
            for i, b in enumerate(batches):
                out[f"critic{st}.{kind}.{i}.share_obs"] = b[0][:, 0]
                out[f"critic{st}.{kind}.{i}.rnn"] = b[1].reshape(b[1].shape[0], -1)[:, 0]
                out[f"critic{st}.{kind}.{i}.value_preds"] = b[2][:, 0]
                out[f"critic{st}.{kind}.{i}.masks"] = b[4][:, 0]
    save("generators_index_maps", **out)


if __name__ == "__main__":
    torch.set_num_threads(1)
    if len(sys.argv) > 1 and sys.argv[1] == "generators":
        case_generators()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoint":
        case_checkpoint()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "activations":  # one HAPPO.update / VCritic.update per activation function
        unusable = []
        for name in ACTIVATIONS:
            try:
                case_single_update([(f"act_{name}", dict(activation_func=name), Discrete(5))])
            except ValueError as e:      # mlp.py:20 nn.init.calculate_gain rejects hardswish and identity: no reference net exists
                unusable.append((name, str(e)))
        case_single_update([("act_tanh_box", dict(activation_func="tanh"), Box(3)), ("act_selu_box", dict(activation_func="selu"), Box(2))])
        print("reference cannot build:", unusable)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "hatrpo":  # regenerate only the HATRPO vectors
        case_hatrpo_parts()
        case_hatrpo_train()
        sys.exit(0)
    case_insert()
    case_gae()
    case_valuenorm()
    case_policy()
    case_single_update()
    case_ha_train()
    case_ma_train()
    case_hatrpo_parts()
    case_hatrpo_train()
    case_generators()
