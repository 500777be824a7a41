"""Minimal observation / action space descriptors.

The hot path only inspects ``__class__.__name__``, ``.shape`` and ``.n`` of a space
(harl/utils/envs_tools.py:15-46, harl/models/base/act.py:24-34), so gym is not needed."""


class Box:
    def __init__(self, low=-float("inf"), high=float("inf"), shape=None, dtype="float32"):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else ()

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape

    def __repr__(self):
        return f"Box{self.shape}"


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n

    def __repr__(self):
        return f"Discrete({self.n})"
