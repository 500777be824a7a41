"""Env-side helpers with the reference's names (harl/utils/envs_tools.py): space shapes, seeding,
vector-env factories.  The factories return *batched* env objects (one object stepping all
n_rollout_threads envs at once), the form the reference itself uses for DexHands
(envs_tools.py:51-54) -- the subprocess-per-env wrapper (harl/envs/env_wrappers.py) is out of scope."""
import os
import random

import numpy as np
import torch


def check(value):
    """numpy -> torch (shares memory), anything else unchanged (envs_tools.py:9-12)."""
    return torch.from_numpy(value) if isinstance(value, np.ndarray) else value


def get_shape_from_obs_space(obs_space):
    kind = obs_space.__class__.__name__
    if kind == "Box":
        return obs_space.shape
    if kind == "list":
        return obs_space
    raise NotImplementedError(kind)


def get_shape_from_act_space(act_space):
    kind = act_space.__class__.__name__
    if kind == "Discrete":
        return 1
    if kind in ("MultiDiscrete", "Box", "MultiBinary"):
        return act_space.shape[0]
    raise NotImplementedError(kind)


def set_seed(args):
    """Seed python / numpy / torch exactly like envs_tools.py:228-237."""
    if not args["seed_specify"]:
        args["seed"] = np.random.randint(1000, 10000)
    s = args["seed"]
    random.seed(s)
    np.random.seed(s)
    os.environ["PYTHONHASHSEED"] = str(s)
    torch.manual_seed(s)
    torch.cuda.manual_seed_all(s)


def _make_env(env_name, seed, n_threads, env_args, device=None):
    from harl_b200.envs import make_batched_env

    return make_batched_env(env_name, seed, n_threads, env_args, device)


def make_train_env(env_name, seed, n_threads, env_args, device=None):
    return _make_env(env_name, seed, n_threads, env_args, device)


def make_eval_env(env_name, seed, n_threads, env_args, device=None):
    return _make_env(env_name, seed * 50000, n_threads, env_args, device)


def get_num_agents(env, env_args, envs):
    return envs.n_agents
