"""Oracle: HAPPO actor update, V-critic update, sequential-agent driver.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  CPU PyTorch fp32 + autograd.

Restates (paths relative to /root/reference):
  * HAPPO.update / HAPPO.train          harl/algorithms/actors/happo.py:28-158
  * VCritic.cal_value_loss/update/train harl/algorithms/critics/v_critic.py:75-200
  * clip_grad_norm_ + Adam              torch (happo.py:93-100, on_policy_base.py:37-42)
  * OnPolicyHARunner.train              harl/runners/on_policy_ha_runner.py:11-130
"""
import math

import numpy as np
import torch

from . import buffers as ob
from . import nets as on


class Adam:
    """torch.optim.Adam single-tensor form (L2 weight decay folded into the gradient)."""

    def __init__(self, params, lr, eps, weight_decay=0.0, betas=(0.9, 0.999)):
        self.params = params  # dict name -> leaf tensor (requires_grad)
        self.lr, self.eps, self.wd, self.betas = lr, eps, weight_decay, betas
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = 0

    @torch.no_grad()
    def step(self, grads):
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1.0 - b1**self.t
        bc2 = 1.0 - b2**self.t
        for k, p in self.params.items():
            g = grads[k]
            if self.wd != 0:
                g = g + self.wd * p
            self.m[k].lerp_(g, 1.0 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1.0 - b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(self.m[k], denom, value=-(self.lr / bc1))


def clip_grads(grads, max_norm, do_clip=True):
    """nn.utils.clip_grad_norm_ (happo.py:93-98) / get_grad_norm (models_tools.py:110-117)."""
    total = torch.sqrt(sum((g.detach() ** 2).sum() for g in grads.values()))
    if do_clip:
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        grads = {k: g * coef for k, g in grads.items()}
    return grads, float(total)


def ppo_loss(logp, old_logp, adv, active, factor, entropy, cfg):
    """happo.py:66-91. Returns (policy_loss, total objective fed to backward, imp_weights)."""
    agg = torch.prod if cfg["action_aggregation"] == "prod" else torch.mean
    imp = agg(torch.exp(logp - old_logp), dim=-1, keepdim=True)
    surr1 = imp * adv
    surr2 = torch.clamp(imp, 1.0 - cfg["clip_param"], 1.0 + cfg["clip_param"]) * adv
    inner = torch.sum(factor * torch.min(surr1, surr2), dim=-1, keepdim=True)
    if cfg["use_policy_active_masks"]:
        pl = (-inner * active).sum() / active.sum()
    else:
        pl = -inner.mean()
    return pl, pl - entropy * cfg["entropy_coef"], imp


def happo_update(p, opt, cfg, head, batch):
    """One HAPPO.update (happo.py:28-102) on a materialised minibatch dict of tensors."""
    logp, ent, _, _ = on.actor_evaluate(
        p, cfg, head, batch["obs"], batch["rnn"], batch["actions"], batch["masks"],
        batch.get("avail"), batch["active"])
    pl, total, imp = ppo_loss(logp, batch["old_logp"], batch["adv"], batch["active"],
                              batch["factor"], ent, cfg)
    names = list(p.keys())
    gs = torch.autograd.grad(total, [p[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(names, gs)}
    raw = {k: g.clone() for k, g in grads.items()}
    grads, gnorm = clip_grads(grads, cfg["max_grad_norm"], cfg["use_max_grad_norm"])
    opt.step(grads)
    return dict(policy_loss=float(pl.detach()), dist_entropy=float(ent.detach()), actor_grad_norm=gnorm,
                ratio=float(imp.detach().mean()), raw_grads=raw)


def huber(e, d):
    """models_tools.py:64-68."""
    a = (e.abs() <= d).float()
    b = (e.abs() > d).float()
    return a * e**2 / 2 + b * d * (e.abs() - d / 2)


def value_loss(values, value_preds, returns_, cfg, vn=None):
    """v_critic.py:75-114 (ValueNorm.update on the batch happens BEFORE normalising it)."""
    clipped = value_preds + (values - value_preds).clamp(-cfg["clip_param"], cfg["clip_param"])
    if vn is not None:
        vn.update(returns_.numpy())
        target = torch.from_numpy(vn.normalize(returns_.numpy()))
    else:
        target = returns_
    e_c = target - clipped
    e_o = target - values
    if cfg["use_huber_loss"]:
        l_c, l_o = huber(e_c, cfg["huber_delta"]), huber(e_o, cfg["huber_delta"])
    else:
        l_c, l_o = e_c**2 / 2, e_o**2 / 2
    loss = torch.max(l_o, l_c) if cfg["use_clipped_value_loss"] else l_o
    return loss.mean()


def critic_update(p, opt, cfg, batch, vn=None):
    """VCritic.update, v_critic.py:116-157."""
    values, _ = on.critic_values(p, cfg, batch["share_obs"], batch["rnn"], batch["masks"])
    vl = value_loss(values, batch["value_preds"], batch["returns"], cfg, vn)
    names = list(p.keys())
    gs = torch.autograd.grad(vl * cfg["value_loss_coef"], [p[k] for k in names])
    grads = dict(zip(names, gs))
    raw = {k: g.clone() for k, g in grads.items()}
    grads, gnorm = clip_grads(grads, cfg["max_grad_norm"], cfg["use_max_grad_norm"])
    opt.step(grads)
    return dict(value_loss=float(vl.detach()), critic_grad_norm=gnorm, raw_grads=raw)


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def actor_minibatches(buf, adv, factor, cfg, perm_fn):
    """Materialise HAPPO.train's minibatches for one epoch (happo.py:129-141 + generators).

    ``buf``: dict of NumPy arrays with the reference buffer attribute names.  ``perm_fn(n)``
    supplies the permutation the reference would draw with torch.randperm(n).
    """
    T, N = buf["actions"].shape[:2]
    nmb = cfg["actor_num_mini_batch"]
    has_avail = buf.get("available_actions") is not None

    def fields(sel_tm1, sel_t):
        b = dict(obs=sel_tm1(buf["obs"]), actions=sel_t(buf["actions"]), masks=sel_tm1(buf["masks"]),
                 active=sel_tm1(buf["active_masks"]), old_logp=sel_t(buf["action_log_probs"]),
                 adv=sel_t(adv), factor=sel_t(factor))
        if has_avail:
            b["avail"] = sel_tm1(buf["available_actions"])
        return b

    if cfg["use_recurrent_policy"]:
        L = cfg["data_chunk_length"]
        for t0, n in ob.recurrent_chunk_indices(perm_fn(T * N // L), T, N, nmb, L):
            tt = (t0[None, :] + np.arange(L)[:, None]).reshape(-1)
            nn_ = np.tile(n, L)
            sel = lambda a: a[tt, nn_]
            b = fields(sel, sel)
            b["rnn"] = buf["rnn_states"][t0, n]
            yield {k: _t(v) for k, v in b.items()}
    elif cfg["use_naive_recurrent_policy"]:
        for ids in ob.naive_recurrent_indices(perm_fn(N), T, N, nmb):
            b = fields(lambda a: a[:T, ids].reshape(T * len(ids), *a.shape[2:]),
                       lambda a: a[:, ids].reshape(T * len(ids), *a.shape[2:]))
            b["rnn"] = buf["rnn_states"][0, ids]
            yield {k: _t(v) for k, v in b.items()}
    else:
        for idx in ob.feed_forward_indices(perm_fn(T * N), T, N, nmb):
            t, n = idx // N, idx % N
            sel = lambda a: a[t, n]
            b = fields(sel, sel)
            b["rnn"] = buf["rnn_states"][t, n]
            yield {k: _t(v) for k, v in b.items()}


def happo_train(p, opt, cfg, head, buf, adv, factor, state_type, perm_fn):
    """HAPPO.train, happo.py:104-158. Returns (train_info, list of per-update infos)."""
    info = dict(policy_loss=0.0, dist_entropy=0.0, actor_grad_norm=0.0, ratio=0.0)
    if np.all(buf["active_masks"][:-1] == 0.0):
        return info, []
    if state_type == "EP":
        adv, _, _ = ob.normalize_advantages(adv, buf["active_masks"][:-1])
    ups = []
    for _ in range(cfg["ppo_epoch"]):
        for batch in actor_minibatches(buf, adv, factor, cfg, perm_fn):
            u = happo_update(p, opt, cfg, head, batch)
            ups.append(u)
            for k in info:
                info[k] += u[k]
    n = cfg["ppo_epoch"] * cfg["actor_num_mini_batch"]
    return {k: v / n for k, v in info.items()}, ups


def logp_sweep(p, cfg, head, buf):
    """on_policy_ha_runner.py:66-83: full-buffer evaluate from rnn_states[0] (time-major rows)."""
    T, N = buf["actions"].shape[:2]
    fl = lambda a: _t(a.reshape(T * N, *a.shape[2:]))
    avail = fl(buf["available_actions"][:-1]) if buf.get("available_actions") is not None else None
    with torch.no_grad():
        lp, _, _, _ = on.actor_evaluate(p, cfg, head, fl(buf["obs"][:-1]), _t(buf["rnn_states"][0]),
                                        fl(buf["actions"]), fl(buf["masks"][:-1]), avail,
                                        fl(buf["active_masks"][:-1]))
    return lp


def factor_update(factor, new_lp, old_lp, cfg):
    """on_policy_ha_runner.py:116-124."""
    agg = torch.prod if cfg["action_aggregation"] == "prod" else torch.mean
    r = agg(torch.exp(new_lp - old_lp), dim=-1).reshape(factor.shape).numpy()
    return (factor * r).astype(np.float32)


def critic_minibatches(cbuf, cfg, perm_fn):
    """v_critic.py:176-188 + on_policy_critic_buffer_ep.py:202-369 (EP) / _fp.py:212-390 (FP).

    FP buffers ([T+1, N, A, ...]) are handled by folding (N, A) into one env-like axis, which
    is what ``_ma_cast`` / the FP feed-forward flatten amount to.
    """
    so, vp, ret, rnn, mk = (cbuf[k] for k in ("share_obs", "value_preds", "returns", "rnn_states_critic", "masks"))
    if vp.ndim == 4:  # FP: [T+1, N, A, .] -> [T+1, N*A, .]
        f = lambda a: a.reshape(a.shape[0], a.shape[1] * a.shape[2], *a.shape[3:])
        so, vp, ret, rnn, mk = f(so), f(vp), f(ret), f(rnn), f(mk)
    T = vp.shape[0] - 1
    N = vp.shape[1]
    nmb = cfg["critic_num_mini_batch"]

    def pack(sel, rnn0):
        return {k: _t(v) for k, v in dict(share_obs=sel(so), value_preds=sel(vp), returns=sel(ret),
                                          masks=sel(mk), rnn=rnn0).items()}

    if cfg["use_recurrent_policy"]:
        L = cfg["data_chunk_length"]
        for t0, n in ob.recurrent_chunk_indices(perm_fn(T * N // L), T, N, nmb, L):
            tt = (t0[None, :] + np.arange(L)[:, None]).reshape(-1)
            nn_ = np.tile(n, L)
            yield pack(lambda a: a[tt, nn_], rnn[t0, n])
    elif cfg["use_naive_recurrent_policy"]:
        for ids in ob.naive_recurrent_indices(perm_fn(N), T, N, nmb):
            yield pack(lambda a: a[:T, ids].reshape(T * len(ids), *a.shape[2:]), rnn[0, ids])
    else:
        for idx in ob.feed_forward_indices(perm_fn(T * N), T, N, nmb):
            t, n = idx // N, idx % N
            yield pack(lambda a: a[t, n], rnn[t, n])


def critic_train(p, opt, cfg, cbuf, vn, perm_fn):
    """VCritic.train, v_critic.py:159-200."""
    info = dict(value_loss=0.0, critic_grad_norm=0.0)
    ups = []
    for _ in range(cfg["critic_epoch"]):
        for batch in critic_minibatches(cbuf, cfg, perm_fn):
            u = critic_update(p, opt, cfg, batch, vn)
            ups.append(u)
            for k in info:
                info[k] += u[k]
    n = cfg["critic_epoch"] * cfg["critic_num_mini_batch"]
    return {k: v / n for k, v in info.items()}, ups


def ha_train(actors, critic, cfg, heads, abufs, cbuf, vn, state_type, agent_order, perm_fn):
    """OnPolicyHARunner.train, on_policy_ha_runner.py:11-130.

    ``actors``: list of (params, Adam); ``critic``: (params, Adam).  Returns per-agent infos,
    critic info, and the factor tensor saved for each agent (what update_factor stored).
    """
    T, N = abufs[0]["actions"].shape[:2]
    factor = np.ones((T, N, 1), np.float32)
    adv = ob.advantages(cbuf["returns"], cbuf["value_preds"], vn)
    if state_type == "FP":
        act = np.stack([b["active_masks"] for b in abufs], axis=2)
        adv, _, _ = ob.normalize_advantages(adv, act[:-1])
    infos, factors = {}, {}
    for a in agent_order:
        factors[a] = factor.copy()
        p, opt = actors[a]
        old_lp = logp_sweep(p, cfg, heads[a], abufs[a])
        adv_a = adv.copy() if state_type == "EP" else adv[:, :, a].copy()
        infos[a], _ = happo_train(p, opt, cfg, heads[a], abufs[a], adv_a, factor, state_type, perm_fn)
        new_lp = logp_sweep(p, cfg, heads[a], abufs[a])
        factor = factor_update(factor, new_lp, old_lp, cfg)
    cinfo, _ = critic_train(critic[0], critic[1], cfg, cbuf, vn, perm_fn)
    return infos, cinfo, factors, factor


# ------------------------------------------------------------------ MAPPO (harl/algorithms/actors/mappo.py)
def mappo_train(p, opt, cfg, head, buf, adv, state_type, perm_fn):
    """MAPPO.train, mappo.py:95-147: HAPPO.train without the importance-ratio factor (factor == 1)."""
    T, N = buf["actions"].shape[:2]
    return happo_train(p, opt, cfg, head, buf, adv, np.ones((T, N, 1), np.float32), state_type, perm_fn)


def mappo_share_param_train(p, opt, cfg, head, abufs, adv, state_type, perm_fn):
    """MAPPO.share_param_train, mappo.py:149-222: one shared actor; every update consumes the CONCATENATION of one
    minibatch per agent; EP advantages are normalised over the agents' stacked active entries."""
    A = len(abufs)
    T, N = abufs[0]["actions"].shape[:2]
    if state_type == "EP":
        stack = np.stack([adv.copy() for _ in range(A)])
        masked = stack.copy()
        for a in range(A):
            masked[a][abufs[a]["active_masks"][:-1] == 0.0] = np.nan
        mean, std = np.nanmean(masked), np.nanstd(masked)
        advs = list((stack - mean) / (std + 1e-5))
    else:
        advs = [adv[:, :, a] for a in range(A)]
    ones = np.ones((T, N, 1), np.float32)
    info = dict(policy_loss=0.0, dist_entropy=0.0, actor_grad_norm=0.0, ratio=0.0)
    for _ in range(cfg["ppo_epoch"]):
        gens = [actor_minibatches(abufs[a], advs[a].astype(np.float32), ones, cfg, perm_fn) for a in range(A)]
        for _ in range(cfg["actor_num_mini_batch"]):
            parts = [next(g) for g in gens]
            batch = {k: torch.cat([b[k] for b in parts], dim=0) for k in parts[0]}
            u = happo_update(p, opt, cfg, head, batch)
            for k in info:
                info[k] += u[k]
    n = cfg["ppo_epoch"] * cfg["actor_num_mini_batch"]
    return {k: v / n for k, v in info.items()}


def ma_train(actors, critic, cfg, heads, abufs, cbuf, vn, state_type, share_param, perm_fn):
    """OnPolicyMARunner.train, on_policy_ma_runner.py:10-60."""
    adv = ob.advantages(cbuf["returns"], cbuf["value_preds"], vn)
    if state_type == "FP":
        act = np.stack([b["active_masks"] for b in abufs], axis=2)
        adv, _, _ = ob.normalize_advantages(adv, act[:-1])
    infos = {}
    if share_param:
        info = mappo_share_param_train(actors[0][0], actors[0][1], cfg, heads[0], abufs, adv.copy(), state_type, perm_fn)
        infos = {a: info for a in range(len(abufs))}
    else:
        for a in range(len(abufs)):
            p, opt = actors[a]
            adv_a = adv.copy() if state_type == "EP" else adv[:, :, a].copy()
            infos[a], _ = mappo_train(p, opt, cfg, heads[a], abufs[a], adv_a, state_type, perm_fn)
    cinfo, _ = critic_train(critic[0], critic[1], cfg, cbuf, vn, perm_fn)
    return infos, cinfo
