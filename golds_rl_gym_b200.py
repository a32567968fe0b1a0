"""Importable alias of the package directory ``golds-rl-gym_b200/`` (a hyphen is not a valid
identifier).  ``import golds_rl_gym_b200 as pkg`` returns the very same module object as
``importlib.import_module("golds-rl-gym_b200")``; submodules are reached as attributes
(``pkg.envs.multiagent``) or through ``submodule("envs.multiagent")``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)

PACKAGE = "golds-rl-gym_b200"
_pkg = importlib.import_module(PACKAGE)


def submodule(name):
    return importlib.import_module(PACKAGE + "." + name)


_pkg.submodule = submodule
sys.modules[__name__] = _pkg
