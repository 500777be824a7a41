"""Oracle: HATRPO actor update (trust-region step of the sequential-agent scheme).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  CPU PyTorch fp32 + autograd (float64 for the
Gaussian KL, as the reference).  Parameter vectors are kept as {name: tensor} dicts; every dot
product runs over all tensors, which is what the reference's flat vectors amount to.

Restates (paths relative to /root/reference):
  * kl_approx / _kl_normal_normal / kl_divergence   harl/utils/trpo_util.py:49-96
  * fisher_vector_product (double backward + 0.1 p)  harl/utils/trpo_util.py:136-158
  * conjugate_gradient (10 steps, residual_tol)      harl/utils/trpo_util.py:100-133
  * HATRPO.update / HATRPO.train                     harl/algorithms/actors/hatrpo.py:37-247
Pinned by tests/golden/hatrpo_*.npz (outputs of the unmodified reference).
"""
import numpy as np
import torch

from . import algo as oa
from . import buffers as ob
from . import nets as on


def _names(p):
    return list(p.keys())


def _dot(a, b):
    return sum((a[k] * b[k]).sum() for k in a)


def kl_divergence(p_new, p_old, cfg, head, batch):
    """trpo_util.py:65-96: KL(old || new) per row, [B, 1]; the old distribution carries no gradient."""
    args = (batch["obs"], batch["rnn"], batch["actions"], batch["masks"], batch.get("avail"), batch["active"])
    _, _, new, _ = on.actor_evaluate(p_new, cfg, head, *args)
    with torch.no_grad():
        _, _, old, _ = on.actor_evaluate(p_old, cfg, head, *args)
    if head == "Discrete":
        pl, ql = old[1], new[1]  # normalised logits (torch Categorical.logits)
        kl = torch.exp(ql - pl) - 1 - ql + pl  # kl_approx, trpo_util.py:49-53
    else:
        pm, ps, qm, qs = (t.to(torch.float64) for t in (old[1], old[2], new[1], new[2]))
        var_ratio = (ps / qs).pow(2)
        t1 = ((pm - qm) / qs).pow(2)
        kl = 0.5 * (var_ratio + t1 - 1 - var_ratio.log())
    return kl.sum(1, keepdim=True)


def fisher_vector_product(p, cfg, head, batch, vec):
    """trpo_util.py:136-158: Hessian-vector product of mean KL(pi || pi) at the current parameters + 0.1 vec."""
    names = _names(p)
    kl = kl_divergence(p, p, cfg, head, batch).mean()
    g = torch.autograd.grad(kl, [p[k] for k in names], create_graph=True, allow_unused=True)
    gv = sum((gi * vec[k]).sum() for k, gi in zip(names, g) if gi is not None)
    h = torch.autograd.grad(gv, [p[k] for k in names], allow_unused=True)
    return {k: ((hi if hi is not None else torch.zeros_like(p[k])).detach().to(torch.float32) + 0.1 * vec[k])
            for k, hi in zip(names, h)}


def conjugate_gradient(p, cfg, head, batch, b, nsteps=10, residual_tol=1e-10):
    """trpo_util.py:100-133."""
    x = {k: torch.zeros_like(v) for k, v in b.items()}
    r = {k: v.clone() for k, v in b.items()}
    d = {k: v.clone() for k, v in b.items()}
    rdotr = _dot(r, r)
    for _ in range(nsteps):
        avp = fisher_vector_product(p, cfg, head, batch, d)
        alpha = rdotr / _dot(d, avp)
        for k in x:
            x[k] = x[k] + alpha * d[k]
            r[k] = r[k] - alpha * avp[k]
        new_rdotr = _dot(r, r)
        beta = new_rdotr / rdotr
        for k in d:
            d[k] = r[k] + beta * d[k]
        rdotr = new_rdotr
        if rdotr < residual_tol:
            break
    return x


def surrogate(p, cfg, head, batch):
    """hatrpo.py:67-92: the maximised objective mean(ratio * factor * adv) (active-masked), plus entropy and ratio."""
    logp, ent, _, _ = on.actor_evaluate(p, cfg, head, batch["obs"], batch["rnn"], batch["actions"], batch["masks"],
                                        batch.get("avail"), batch["active"])
    agg = torch.prod if cfg["action_aggregation"] == "prod" else torch.mean
    ratio = agg(torch.exp(logp - batch["old_logp"]), dim=-1, keepdim=True)
    inner = torch.sum(ratio * batch["factor"] * batch["adv"], dim=-1, keepdim=True)
    if cfg["use_policy_active_masks"]:
        loss = (inner * batch["active"]).sum() / batch["active"].sum()
    else:
        loss = inner.mean()
    return loss, ent, ratio


def hatrpo_update(p, cfg, head, batch):
    """One HATRPO.update (hatrpo.py:37-194).  Mutates ``p`` in place; returns the reference's five outputs plus the
    pieces the parity tests pin (gradient, step direction, accepted fraction)."""
    names = _names(p)
    loss, _, _ = surrogate(p, cfg, head, batch)
    gs = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    g = {k: (gi if gi is not None else torch.zeros_like(p[k])).detach() for k, gi in zip(names, gs)}
    step_dir = conjugate_gradient(p, cfg, head, batch, g, nsteps=10)
    loss0 = float(loss.detach())
    fvp = fisher_vector_product(p, cfg, head, batch, step_dir)
    shs = 0.5 * _dot(step_dir, fvp)
    step_size = 1.0 / torch.sqrt(shs / cfg["kl_threshold"])
    full = {k: step_size * v for k, v in step_dir.items()}
    old = {k: v.detach().clone() for k, v in p.items()}
    expected = float(_dot(g, full))
    flag, fraction, accepted = False, 1.0, None
    kl = ent = ratio = None
    improve = 0.0
    for i in range(cfg["ls_step"]):
        with torch.no_grad():
            for k in names:
                p[k].copy_(old[k] + fraction * full[k])
        with torch.no_grad():
            new_loss, ent, ratio = surrogate(p, cfg, head, batch)
            kl = kl_divergence(p, old, cfg, head, batch).mean()
        improve = float(np.float32(new_loss.item()) - np.float32(loss0))
        if float(kl) < cfg["kl_threshold"] and improve / expected > cfg["accept_ratio"] and improve > 0:
            flag, accepted = True, i
            break
        expected *= cfg["backtrack_coeff"]
        fraction *= cfg["backtrack_coeff"]
    if not flag:
        with torch.no_grad():
            for k in names:
                p[k].copy_(old[k])
    return dict(kl=float(kl), loss_improve=improve, expected_improve=expected, dist_entropy=float(ent),
                ratio=float(ratio.mean()), loss=loss0, grad=g, step_dir=step_dir, accepted=accepted)


def hatrpo_train(p, cfg, head, buf, adv, factor, state_type, perm_fn):
    """HATRPO.train, hatrpo.py:196-247: one pass, one minibatch = the whole buffer."""
    info = dict(kl=0.0, dist_entropy=0.0, loss_improve=0.0, expected_improve=0.0, ratio=0.0)
    if np.all(buf["active_masks"][:-1] == 0.0):
        return info
    if state_type == "EP":
        adv, _, _ = ob.normalize_advantages(adv, buf["active_masks"][:-1])
    cfg1 = dict(cfg, actor_num_mini_batch=1)
    for batch in oa.actor_minibatches(buf, adv, factor, cfg1, perm_fn):
        u = hatrpo_update(p, cfg, head, batch)
        for k in info:
            info[k] += u[k]
    return info


def ha_train_hatrpo(actors, critic, cfg, heads, abufs, cbuf, vn, state_type, agent_order, perm_fn):
    """OnPolicyHARunner.train (on_policy_ha_runner.py:11-130) with HATRPO actors (``actors``: list of param dicts)."""
    T, N = abufs[0]["actions"].shape[:2]
    factor = np.ones((T, N, 1), np.float32)
    adv = ob.advantages(cbuf["returns"], cbuf["value_preds"], vn)
    if state_type == "FP":
        act = np.stack([b["active_masks"] for b in abufs], axis=2)
        adv, _, _ = ob.normalize_advantages(adv, act[:-1])
    infos, factors = {}, {}
    for a in agent_order:
        factors[a] = factor.copy()
        p = actors[a]
        old_lp = oa.logp_sweep(p, cfg, heads[a], abufs[a])
        adv_a = adv.copy() if state_type == "EP" else adv[:, :, a].copy()
        infos[a] = hatrpo_train(p, cfg, heads[a], abufs[a], adv_a, factor, state_type, perm_fn)
        new_lp = oa.logp_sweep(p, cfg, heads[a], abufs[a])
        factor = oa.factor_update(factor, new_lp, old_lp, cfg)
    cinfo, _ = oa.critic_train(critic[0], critic[1], cfg, cbuf, vn, perm_fn)
    return infos, cinfo, factors, factor


def fisher_vector_product_gn(p, cfg, head, batch, vec):
    """The same product in Gauss-Newton form J^T H (J vec) / B + 0.1 vec -- the form the CUDA path evaluates.

    At new == old the gradient of the KL w.r.t. the distribution parameters vanishes, so the Hessian of
    trpo_util.py:136-158 reduces to J^T H J with H = d2 KL / d(dist params)^2: the identity over ALL normalised
    logits (kl_approx, masked entries included) for Categorical; diag(1/sigma^2) over the means and 2 over
    log sigma for the Gaussian.  Checked against the reference's double-backward golden vectors."""
    names = _names(p)
    args = (batch["obs"], batch["rnn"], batch["actions"], batch["masks"], batch.get("avail"), batch["active"])

    def dist_params(*flat):
        q = dict(zip(names, flat))
        _, _, d, _ = on.actor_evaluate(q, cfg, head, *args)
        if head == "Discrete":
            return (d[1],)
        return d[1], d[2].log()

    prim = tuple(p[k].detach() for k in names)
    tang = tuple(vec[k] for k in names)
    outs, jv = torch.autograd.functional.jvp(dist_params, prim, tang)
    B = outs[0].shape[0]
    if head == "Discrete":
        cot = (jv[0] / B,)
    else:
        sig2 = torch.exp(2 * outs[1])
        cot = (jv[0] / sig2 / B, 2.0 * jv[1] / B)
    _, hv = torch.autograd.functional.vjp(dist_params, prim, cot)
    return {k: h + 0.1 * vec[k] for k, h in zip(names, hv)}
