"""Importable alias for the package directory `opencv-octvr_b200/` (hyphenated, per the repo layout)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("opencv-octvr_b200")
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
capi = importlib.import_module("opencv-octvr_b200.capi")
