"""Actor registry (reference: harl/algorithms/actors/__init__.py:13-24); on-policy algorithms only."""
from .haa2c import HAA2C
from .happo import HAPPO
from .hatrpo import HATRPO
from .mappo import MAPPO

ALGO_REGISTRY = {"happo": HAPPO, "hatrpo": HATRPO, "haa2c": HAA2C, "mappo": MAPPO}
