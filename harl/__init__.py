"""``harl`` import alias so that the reference's entry points run unchanged against harl_b200.

``examples/train.py`` of the reference imports ``harl.utils.configs_tools`` and
``harl.runners.RUNNER_REGISTRY`` (train.py:4,87).  This package maps every ``harl.x.y`` import to
the ``harl_b200.x.y`` module object (one module, two names), so user code written against the
reference package layout keeps working.
"""
import importlib
import importlib.abc
import importlib.util
import sys

import harl_b200


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("harl."):
            return None
        real = "harl_b200." + fullname[len("harl."):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, _AliasLoader(real))


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
__path__ = []
__version__ = harl_b200.__version__
