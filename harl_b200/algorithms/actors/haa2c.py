"""HAA2C = HAPPO without ratio clipping (reference: harl/algorithms/actors/haa2c.py; `a2c_epoch` key)."""
from .happo import HAPPO


class HAA2C(HAPPO):
    use_clip = False
