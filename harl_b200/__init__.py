"""harl_b200: B200-native implementation of HARL's on-policy HAPPO hot path.

Python keeps the reference's Runner / Algorithm / Buffer / Env surface; every tensor op on the
path is a hand-written sm_100a CUDA kernel behind the C ABI in include/harl_b200.h.
"""
__version__ = "0.1.0"
