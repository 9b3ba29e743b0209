"""Import the *unmodified* reference modules for this path inside the build
container (TEST INFRASTRUCTURE ONLY; ``/root/reference`` does not exist on the
GPU box, so nothing that runs there may call :func:`load_reference`).

A plain ``import calibration.WATS`` fails here because the package
``__init__`` pulls matplotlib (calibration/__init__.py:20 -> TS.py:17) and
``utils/ece.py:3,6`` imports matplotlib/seaborn at top level.  We register
bare package modules (so the ``__init__`` files never run) and empty stubs for
the plotting libraries, then import the three files the path needs.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EGNN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "calibration", "WATS.py"))


def load_reference():
    """Returns ``(wats_module, model_module, ece_module)`` of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])
    for pkg, rel in (("calibration", "calibration"), ("src", "src"),
                     ("src.gnn", os.path.join("src", "gnn")), ("utils", "utils")):
        if pkg not in sys.modules or not getattr(sys.modules[pkg], "__egnn_shim__", False):
            mod = types.ModuleType(pkg)
            mod.__path__ = [os.path.join(REFERENCE_ROOT, rel)]
            mod.__egnn_shim__ = True
            sys.modules[pkg] = mod
    wats = importlib.import_module("calibration.WATS")
    model = importlib.import_module("src.gnn.model")
    ece = importlib.import_module("utils.ece")
    return wats, model, ece
