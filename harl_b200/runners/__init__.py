"""Runner registry (reference: harl/runners/__init__.py:7-18); the on-policy HA path only."""
from .on_policy_ha_runner import OnPolicyHARunner

RUNNER_REGISTRY = {"happo": OnPolicyHARunner, "haa2c": OnPolicyHARunner}
