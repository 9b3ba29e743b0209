"""Import alias for the package directory ``efficient-gnn_b200/``.

The product lives in ``efficient-gnn_b200/`` (the name the repo layout
prescribes); a hyphen is not importable, so this shim makes
``import efficient_gnn_b200`` resolve every submodule from that directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "efficient-gnn_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
