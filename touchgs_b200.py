"""Importable alias of the ``touch-gs_b200/`` package directory (hyphen in its name)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_real = importlib.import_module("touch-gs_b200")
sys.modules[__name__] = _real
