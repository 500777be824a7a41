"""Runner registry (reference: harl/runners/__init__.py:7-18); the on-policy path only."""
from .on_policy_ha_runner import OnPolicyHARunner
from .on_policy_ma_runner import OnPolicyMARunner

RUNNER_REGISTRY = {"happo": OnPolicyHARunner, "hatrpo": OnPolicyHARunner, "haa2c": OnPolicyHARunner,
                   "mappo": OnPolicyMARunner}
