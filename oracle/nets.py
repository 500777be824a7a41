"""Oracle: actor / critic networks as pure functions over a ``{state_dict key: tensor}`` dict.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  CPU PyTorch fp32.

Restates (reference paths relative to /root/reference):
  * MLPBase / MLPLayer          harl/models/base/mlp.py:7-70
  * RNNLayer                    harl/models/base/rnn.py:8-81
  * ACTLayer                    harl/models/base/act.py:44-80,143-157
  * Categorical / DiagGaussian  harl/models/base/distributions.py:7-89
  * StochasticPolicy            harl/models/policy_models/stochastic_policy.py:55-127
  * VNet                        harl/models/value_function_models/v_net.py:48-67
Parameter names are the reference ``state_dict`` keys, so a reference checkpoint is a
valid parameter dict here.
"""
import math

import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


def activation(name):
    """harl/utils/models_tools.py:28-50 (get_active_func)."""
    return {
        "sigmoid": torch.sigmoid,
        "tanh": torch.tanh,
        "relu": torch.relu,
        "leaky_relu": lambda x: F.leaky_relu(x, 0.01),
        "selu": F.selu,
        "hardswish": F.hardswish,
        "identity": lambda x: x,
    }[name]


def n_hidden_layers(cfg):
    return len(cfg["hidden_sizes"])


def uses_rnn(cfg):
    return bool(cfg["use_recurrent_policy"] or cfg["use_naive_recurrent_policy"])


def init_params(cfg, in_dim, head, out_dim, generator=None):
    """Fresh parameters with the reference's initialisation scheme.

    mlp.py:17-23 (orthogonal, gain=calculate_gain(act), bias 0), rnn.py:15-20,
    distributions.py:43-49,74-82 (gain=cfg['gain']), v_net.py:41-44 (gain 1).
    ``head`` in {"Discrete", "Box", "value"}.
    """
    g = generator
    init = getattr(torch.nn.init, cfg["initialization_method"])
    gain = torch.nn.init.calculate_gain(cfg["activation_func"])
    p = {}

    def lin(prefix, o, i, gn):
        w = torch.empty(o, i)
        init(w, gain=gn, generator=g) if g is not None else init(w, gain=gn)
        p[prefix + ".weight"] = w
        p[prefix + ".bias"] = torch.zeros(o)

    def ln(prefix, d):
        p[prefix + ".weight"] = torch.ones(d)
        p[prefix + ".bias"] = torch.zeros(d)

    if cfg["use_feature_normalization"]:
        ln("base.feature_norm", in_dim)
    prev = in_dim
    for li, h in enumerate(cfg["hidden_sizes"]):
        lin(f"base.mlp.fc.{3 * li}", h, prev, gain)
        ln(f"base.mlp.fc.{3 * li + 2}", h)
        prev = h
    if uses_rnn(cfg):
        for l in range(cfg["recurrent_n"]):
            for nm in ("weight_ih", "weight_hh"):
                w = torch.empty(3 * prev, prev)
                init(w, generator=g) if g is not None else init(w)
                p[f"rnn.rnn.{nm}_l{l}"] = w
            p[f"rnn.rnn.bias_ih_l{l}"] = torch.zeros(3 * prev)
            p[f"rnn.rnn.bias_hh_l{l}"] = torch.zeros(3 * prev)
        ln("rnn.norm", prev)
    if head == "Discrete":
        lin("act.action_out.linear", out_dim, prev, cfg["gain"])
    elif head == "Box":
        p["act.action_out.log_std"] = torch.ones(out_dim) * cfg["std_x_coef"]
        lin("act.action_out.fc_mean", out_dim, prev, cfg["gain"])
    elif head == "value":
        lin("v_out", 1, prev, 1.0)
    else:
        raise NotImplementedError(head)
    return p


def trunk(p, cfg, x):
    """MLPBase.forward, mlp.py:64-70: LN(obs) -> [Linear -> act -> LN] * L."""
    act = activation(cfg["activation_func"])
    if cfg["use_feature_normalization"]:
        x = F.layer_norm(x, x.shape[-1:], p["base.feature_norm.weight"], p["base.feature_norm.bias"])
    for li in range(n_hidden_layers(cfg)):
        x = F.linear(x, p[f"base.mlp.fc.{3 * li}.weight"], p[f"base.mlp.fc.{3 * li}.bias"])
        x = act(x)
        x = F.layer_norm(x, x.shape[-1:], p[f"base.mlp.fc.{3 * li + 2}.weight"], p[f"base.mlp.fc.{3 * li + 2}.bias"])
    return x


def gru_cell(p, layer, x, h):
    """One torch.nn.GRU layer step, gate order (r, z, n); SURVEY Appendix A."""
    gi = F.linear(x, p[f"rnn.rnn.weight_ih_l{layer}"], p[f"rnn.rnn.bias_ih_l{layer}"])
    gh = F.linear(h, p[f"rnn.rnn.weight_hh_l{layer}"], p[f"rnn.rnn.bias_hh_l{layer}"])
    i_r, i_z, i_n = gi.chunk(3, dim=-1)
    h_r, h_z, h_n = gh.chunk(3, dim=-1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1.0 - z) * n + z * h


def rnn_layer(p, cfg, x, hxs, masks):
    """RNNLayer.forward, rnn.py:23-81.

    ``x`` [B, h]; ``hxs`` [N, R, h]; ``masks`` [B, 1].  If B == N: one step.  Else B = T*N,
    time-major, and the GRU runs over T steps from ``hxs`` with ``h <- h * mask_t`` before
    each step (the reference's segment splitting at reset steps is an optimisation of
    exactly this recurrence, rnn.py:46-70).  Returns (LN(out) [B, h], hxs' [N, R, h]).
    """
    R = cfg["recurrent_n"]
    N = hxs.shape[0]
    B = x.shape[0]
    T = B // N
    xs = x.view(T, N, -1)
    ms = masks.view(T, N, 1)
    h = [hxs[:, l] for l in range(R)]
    outs = []
    for t in range(T):
        inp = xs[t]
        for l in range(R):
            h[l] = gru_cell(p, l, inp, h[l] * ms[t])
            inp = h[l]
        outs.append(inp)
    out = torch.stack(outs, 0).reshape(B, -1)
    out = F.layer_norm(out, out.shape[-1:], p["rnn.norm.weight"], p["rnn.norm.bias"])
    return out, torch.stack(h, dim=1)


def features(p, cfg, x, hxs, masks):
    f = trunk(p, cfg, x)
    if uses_rnn(cfg):
        f, hxs = rnn_layer(p, cfg, f, hxs, masks)
    return f, hxs


def categorical_logits(p, feat, avail):
    """Categorical.forward, distributions.py:51-55; normalised like torch Categorical(logits=)."""
    x = F.linear(feat, p["act.action_out.linear.weight"], p["act.action_out.linear.bias"])
    if avail is not None:
        x = torch.where(avail == 0, torch.full_like(x, -1e10), x)
    return x - x.logsumexp(dim=-1, keepdim=True)


def gaussian_params(p, cfg, feat):
    """DiagGaussian.forward, distributions.py:86-89."""
    mean = F.linear(feat, p["act.action_out.fc_mean.weight"], p["act.action_out.fc_mean.bias"])
    std = torch.sigmoid(p["act.action_out.log_std"] / cfg["std_x_coef"]) * cfg["std_y_coef"]
    return mean, std.expand_as(mean)


def actor_evaluate(p, cfg, head, obs, hxs, action, masks, avail=None, active=None):
    """StochasticPolicy.evaluate_actions, stochastic_policy.py:93-127 + act.py:143-157.

    Returns (log_probs [B, 1] or [B, act_dim], entropy scalar, dist) where dist is
    ("Discrete", normalised logits) or ("Box", mean, std); plus the new hidden state.
    Entropy is the active-masked mean iff cfg['use_policy_active_masks'] and ``active``
    is given (stochastic_policy.py:124), else the plain mean.
    """
    feat, hxs = features(p, cfg, obs, hxs, masks)
    if head == "Discrete":
        logits = categorical_logits(p, feat, avail)
        logp = logits.gather(-1, action.long().view(-1, 1))
        probs = logits.exp()
        # torch.distributions.Categorical.entropy clamps logits at finfo.min first
        ent = -(torch.clamp(logits, min=torch.finfo(logits.dtype).min) * probs).sum(-1)
        dist = ("Discrete", logits)
    else:
        mean, std = gaussian_params(p, cfg, feat)
        var = std * std
        logp = -((action - mean) ** 2) / (2 * var) - std.log() - 0.5 * LOG_2PI
        ent = (0.5 + 0.5 * LOG_2PI + std.log()).sum(-1)
        dist = ("Box", mean, std)
    if active is not None and cfg["use_policy_active_masks"]:
        entropy = (ent * active.squeeze(-1)).sum() / active.sum()
    else:
        entropy = ent.mean()
    return logp, entropy, dist, hxs


def actor_mode(p, cfg, head, obs, hxs, masks, avail=None):
    """StochasticPolicy.forward with deterministic=True (act.py:72-80): (action, logp, hxs)."""
    feat, hxs = features(p, cfg, obs, hxs, masks)
    if head == "Discrete":
        logits = categorical_logits(p, feat, avail)
        a = logits.argmax(-1, keepdim=True)
        return a.float(), logits.gather(-1, a), hxs
    mean, std = gaussian_params(p, cfg, feat)
    return mean, (-std.log() - 0.5 * LOG_2PI).expand_as(mean), hxs


def critic_values(p, cfg, cent_obs, hxs, masks):
    """VNet.forward, v_net.py:48-67."""
    feat, hxs = features(p, cfg, cent_obs, hxs, masks)
    return F.linear(feat, p["v_out.weight"], p["v_out.bias"]), hxs
