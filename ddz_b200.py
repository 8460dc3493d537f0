"""Importable alias of the package directory `doudizhu-rl_b200/` (a hyphen is not a valid Python identifier)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("doudizhu-rl_b200")
sys.modules[__name__] = _pkg
