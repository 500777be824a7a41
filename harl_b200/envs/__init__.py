"""Environment side of the L0 boundary: batched env factory and logger registry
(reference: harl/envs/__init__.py:14-23, harl/utils/envs_tools.py:49-107)."""
from .logger import OnPolicyLogger
from .spaces import Box, Discrete  # noqa: F401
from .synthetic import SyntheticBatchedEnv

ENV_NAMES = ("smac", "mamujoco", "pettingzoo_mpe", "gym", "football", "dexhands", "smacv2", "lag", "synthetic")
LOGGER_REGISTRY = {name: OnPolicyLogger for name in ENV_NAMES}


def make_batched_env(env_name, seed, n_threads, env_args, device=None):
    """One env object stepping all ``n_threads`` environments at once, tensors on ``device``.

    ``pettingzoo_mpe`` simple_spread has a native batched implementation (``env_args['backend']``
    = "native", default when available); every other name -- the third-party simulators are not in
    this image -- is served by the seeded synthetic env with that task's tensor shapes."""
    if env_name == "pettingzoo_mpe" and env_args.get("backend", "synthetic") == "native":
        from .mpe_spread import BatchedSimpleSpread

        return BatchedSimpleSpread(seed, n_threads, env_args, device)
    if env_name not in ENV_NAMES:
        raise NotImplementedError(f"Can not support the {env_name} environment.")
    return SyntheticBatchedEnv(env_name, seed, n_threads, env_args, device)
