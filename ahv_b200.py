"""Importable alias for the package directory `3dahv_b200/` (not a valid identifier)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
_pkg = importlib.import_module("3dahv_b200")
sys.modules[__name__] = _pkg
