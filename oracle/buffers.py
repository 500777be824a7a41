"""Oracle: rollout-buffer logic (masks, GAE/returns, advantages, minibatch index maps).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Strict-fp32 NumPy; the arithmetic order
follows the reference expression by expression so results are bit-identical to it.

Restates (paths relative to /root/reference):
  * OnPolicyBaseRunner.insert mask derivation   harl/runners/on_policy_base_runner.py:342-433
  * OnPolicyCriticBuffer{EP,FP}.compute_returns  harl/common/buffers/on_policy_critic_buffer_ep.py:97-200
                                                 harl/common/buffers/on_policy_critic_buffer_fp.py:107-210
  * advantages + masked normalisation            harl/runners/on_policy_ha_runner.py:26-45,
                                                 harl/algorithms/actors/happo.py:122-127
  * ValueNorm                                    harl/common/valuenorm.py:38-92
  * minibatch index maps                         harl/common/buffers/on_policy_actor_buffer.py:114-326
"""
import numpy as np

f32 = np.float32


# ---------------------------------------------------------------- insert: mask derivation
def derive_masks(dones, bad_transition, state_type="EP"):
    """on_policy_base_runner.py:358-433.

    dones [N, A] bool; bad_transition [N, A] bool (``infos[n][a]["bad_transition"]``).
    Returns masks [N, A, 1], active_masks [N, A, 1], bad_masks ([N, 1] EP from agent 0 /
    [N, A, 1] FP), dones_env [N] -- all fp32 {0, 1}.
    """
    dones = np.asarray(dones, dtype=bool)
    N, A = dones.shape
    dones_env = np.all(dones, axis=1)
    masks = np.ones((N, A, 1), f32)
    masks[dones_env] = 0.0
    active = np.ones((N, A, 1), f32)
    active[dones] = 0.0
    active[dones_env] = 1.0
    bt = np.asarray(bad_transition, dtype=bool)
    if state_type == "EP":
        bad = np.where(bt[:, 0:1], f32(0.0), f32(1.0)).astype(f32)  # [N, 1]
    else:
        bad = np.where(bt[:, :, None], f32(0.0), f32(1.0)).astype(f32)  # [N, A, 1]
    return masks, active, bad, dones_env


# ---------------------------------------------------------------- ValueNorm
class ValueNormState:
    """harl/common/valuenorm.py:7-92 with input_shape=1, per_element_update=False."""

    def __init__(self, beta=0.99999, epsilon=1e-5):
        self.beta = beta
        self.epsilon = epsilon
        self.running_mean = f32(0.0)
        self.running_mean_sq = f32(0.0)
        self.debiasing_term = f32(0.0)

    def mean_var(self):
        """valuenorm.py:38-45."""
        d = max(self.debiasing_term, f32(self.epsilon))
        m = f32(self.running_mean / d)
        msq = f32(self.running_mean_sq / d)
        var = max(f32(msq - f32(m * m)), f32(1e-2))
        return f32(m), f32(var)

    def update(self, x):
        """valuenorm.py:47-64 (fp32 torch mean over the batch, then EMA)."""
        import torch

        t = torch.as_tensor(np.asarray(x, dtype=f32)).reshape(-1)
        bm = f32(t.mean().item())
        bsq = f32((t**2).mean().item())
        w = f32(self.beta)
        omw = f32(1.0 - self.beta)  # python double -> fp32 scalar multiply in torch
        self.running_mean = f32(f32(self.running_mean * w) + f32(bm * omw))
        self.running_mean_sq = f32(f32(self.running_mean_sq * w) + f32(bsq * omw))
        self.debiasing_term = f32(f32(self.debiasing_term * w) + f32(f32(1.0) * omw))

    def normalize(self, x):
        m, v = self.mean_var()
        return ((np.asarray(x, f32) - m) / f32(np.sqrt(v))).astype(f32)

    def denormalize(self, x):
        m, v = self.mean_var()
        return (np.asarray(x, f32) * f32(np.sqrt(v)) + m).astype(f32)


# ---------------------------------------------------------------- GAE / returns
def compute_returns(rewards, value_preds, masks, bad_masks, next_value, gamma, gae_lambda,
                    use_gae=True, use_proper_time_limits=True, vn=None):
    """on_policy_critic_buffer_ep.py:97-200 (the FP twin differs only in array rank).

    rewards [T, ...], value_preds / masks / bad_masks [T+1, ...] fp32, next_value [...].
    ``vn`` is a ValueNormState or None.  Returns (returns [T+1, ...], value_preds with
    slot T overwritten when use_gae) -- copies, inputs untouched.  NumPy weak-scalar rule:
    ``gamma`` and ``gamma * gae_lambda`` are Python doubles cast to fp32 when they meet
    an fp32 array, so the product gamma*lambda is formed in double first.
    """
    rewards = np.asarray(rewards, f32)
    vp = np.array(value_preds, f32, copy=True)
    masks = np.asarray(masks, f32)
    bad = np.asarray(bad_masks, f32)
    T = rewards.shape[0]
    ret = np.zeros_like(vp)
    den = (lambda a: vn.denormalize(a)) if vn is not None else (lambda a: a)
    g32 = f32(gamma)
    gl32 = f32(gamma * gae_lambda)
    if use_gae:
        vp[-1] = next_value
        gae = np.zeros_like(rewards[0])
        for t in reversed(range(T)):
            delta = rewards[t] + g32 * den(vp[t + 1]) * masks[t + 1] - den(vp[t])
            gae = delta + gl32 * masks[t + 1] * gae
            if use_proper_time_limits:
                gae = bad[t + 1] * gae
            ret[t] = gae + den(vp[t])
    else:
        ret[-1] = next_value
        for t in reversed(range(T)):
            if use_proper_time_limits:
                ret[t] = (ret[t + 1] * g32 * masks[t + 1] + rewards[t]) * bad[t + 1] + (
                    f32(1) - bad[t + 1]
                ) * den(vp[t])
            else:
                ret[t] = ret[t + 1] * g32 * masks[t + 1] + rewards[t]
    return ret, vp


def advantages(returns, value_preds, vn=None):
    """on_policy_ha_runner.py:26-33: returns[:-1] - denorm(value_preds[:-1])."""
    v = value_preds[:-1]
    if vn is not None:
        v = vn.denormalize(v)
    return (returns[:-1] - v).astype(f32)


def normalize_advantages(adv, active_masks_T):
    """happo.py:122-127 / on_policy_ha_runner.py:36-45: masked nanmean / nanstd, eps 1e-5.

    adv and active_masks_T share a shape ([T, N, 1] EP per agent, [T, N, A, 1] FP global).
    """
    c = adv.copy()
    c[active_masks_T == 0.0] = np.nan
    mean = np.nanmean(c)
    std = np.nanstd(c)
    return ((adv - mean) / (std + 1e-5)).astype(f32), mean, std


# ---------------------------------------------------------------- minibatch index maps
def feed_forward_indices(perm, T, N, num_mini_batch):
    """on_policy_actor_buffer.py:121-137: minibatch i = perm[i*mb:(i+1)*mb], row k -> (k//N, k%N)."""
    mb = (T * N) // num_mini_batch
    return [np.asarray(perm[i * mb:(i + 1) * mb]) for i in range(num_mini_batch)]


def naive_recurrent_indices(perm, T, N, num_mini_batch):
    """on_policy_actor_buffer.py:187-205: env ids per batch; row t*k+j <-> (t, ids[j])."""
    k = N // num_mini_batch
    return [np.asarray(perm[i * k:(i + 1) * k]) for i in range(num_mini_batch)]


def recurrent_chunk_indices(perm, T, N, num_mini_batch, L):
    """on_policy_actor_buffer.py:231-246,266-279: chunk c -> (n = c // (T/L), t0 = (c % (T/L)) * L).

    Returns a list of (t0 [mb], n [mb]) per minibatch; batch row l*mb+j <-> (t0[j]+l, n[j]).
    """
    chunks = (T * N) // L
    mb = chunks // num_mini_batch
    per_env = T // L
    out = []
    for i in range(num_mini_batch):
        c = np.asarray(perm[i * mb:(i + 1) * mb])
        out.append(((c % per_env) * L, c // per_env))
    return out
