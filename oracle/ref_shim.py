"""Import the *unmodified* reference modules for this path (TEST
INFRASTRUCTURE ONLY: ``tests/``, ``smoke()`` and the CPU legs of ``bench.py``).

Two places can hold them:

* ``/root/reference`` - the reference checkout, mounted in the build container
  only;
* ``oracle/_ref/`` - the same modules byte-compiled by ``oracle/build_ref.py``
  (``.pyc`` bytes in ``.bin`` files - the snapshot drops ``*.pyc`` - git-ignored,
  travels to the GPU box with the snapshot).

A plain ``import calibration.WATS`` fails because the package ``__init__``
pulls matplotlib (calibration/__init__.py:20 -> TS.py:17) and
``utils/ece.py:3,6`` imports matplotlib/seaborn at top level.  We register
bare package modules (so the ``__init__`` files never run) and empty stubs for
the plotting libraries, then import the three files the path needs.
"""
from __future__ import annotations

import importlib
import marshal
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EGNN_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _has(root: str, ext: str) -> bool:
    return os.path.isfile(os.path.join(root, "calibration", "WATS" + ext))


def reference_available() -> bool:
    """The reference checkout itself is mounted (build container)."""
    return _has(REFERENCE_ROOT, ".py")


def staged_available() -> bool:
    """``oracle/_ref`` holds the byte-compiled reference modules."""
    return _has(STAGED_ROOT, ".bin")


def reference_root():
    """Where :func:`load_reference` will import from, or None."""
    if reference_available():
        return REFERENCE_ROOT
    if staged_available():
        return STAGED_ROOT
    return None


def load_reference(root=None):
    """Returns ``(wats_module, model_module, ece_module)`` of the reference."""
    root = root or reference_root()
    if root is None:
        raise RuntimeError(f"reference neither mounted at {REFERENCE_ROOT} nor staged in {STAGED_ROOT} "
                           "(run __graft_entry__.build() where /root/reference exists)")
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
        have = sys.modules.get(pkg)
        if have is None or getattr(have, "__egnn_shim__", None) != root:
            for name in [m for m in sys.modules if m == pkg or m.startswith(pkg + ".")]:
                del sys.modules[name]                   # stale import from the other root
            mod = types.ModuleType(pkg)
            mod.__path__ = [os.path.join(root, rel)]
            mod.__egnn_shim__ = root
            sys.modules[pkg] = mod
    importlib.invalidate_caches()
    if root == STAGED_ROOT:
        # byte-compiled modules: unmarshal the code object and run it in a module registered
        # under the reference's own dotted name (dependencies first, so that the relative
        # import in calibration/WATS.py:7 finds calibration.utils in sys.modules)
        for name in ("calibration.utils", "calibration.WATS", "src.gnn.model", "utils.ece"):
            if name in sys.modules and getattr(sys.modules[name], "__egnn_shim__", None) == root:
                continue
            path = os.path.join(root, *name.split(".")) + ".bin"
            with open(path, "rb") as fh:
                code = marshal.loads(fh.read()[16:])           # 16-byte pyc header, then the code object
            mod = types.ModuleType(name)
            mod.__file__ = path
            mod.__package__ = name.rpartition(".")[0]
            mod.__egnn_shim__ = root
            sys.modules[name] = mod
            exec(code, mod.__dict__)
            setattr(sys.modules[mod.__package__], name.rpartition(".")[2], mod)
        return sys.modules["calibration.WATS"], sys.modules["src.gnn.model"], sys.modules["utils.ece"]
    wats = importlib.import_module("calibration.WATS")
    model = importlib.import_module("src.gnn.model")
    ece = importlib.import_module("utils.ece")
    return wats, model, ece
