"""Shared helpers for the test-suite (golden loading, config reconstruction)."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def base_args(**over):
    """Mirror of tests/golden/make_golden.py:base_args (the reference's config keys)."""
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


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def cfg_of(g):
    over = {str(k): ast.literal_eval(str(v)) for k, v in zip(g["cfg_keys"], g["cfg_vals"])}
    return base_args(**over)


def meta_of(g):
    tag, kind, st, A, od, sd, adim = (str(x) for x in g["meta"])
    return dict(tag=tag, head=kind, state_type=st, A=int(A), od=int(od), sd=int(sd), act_dim=int(adim))


def params_of(g, prefix, grad=False):
    p = {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith(prefix)}
    if grad:
        for v in p.values():
            v.requires_grad_(True)
    return p


def sub(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def perm_replayer(g):
    perms = [g[f"perm{i}"] for i in range(int(g["n_perms"]))]
    it = iter(perms)

    def fn(n):
        p = next(it)
        assert len(p) == n, (len(p), n)
        return p

    return fn
