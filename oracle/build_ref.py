"""Recipe for ``oracle/_ref/``: byte-compile the reference's own modules for
this path from the sources where they lie under ``/root/reference``.

TEST INFRASTRUCTURE ONLY (like everything under ``oracle/``).

The reference is pure Python, so its "build" is ``py_compile``: each module is
compiled, unmodified, straight from ``/root/reference/<path>.py`` into
``oracle/_ref/<path>.bin`` - the bytes of a ``.pyc`` file under another
extension, because the snapshot that travels to the GPU box leaves ``*.pyc``
behind (probed: ``.bin`` travels, like the built ``.so``).  No reference source
enters the repo; ``oracle/_ref/`` is git-ignored but not gpurun-ignored, so the
compiled modules reach the GPU box (same image, same interpreter) where
``/root/reference`` does not exist.  ``oracle/ref_shim.py`` loads them there
(marshal -> code object -> module) so that

* ``tests/test_gpu_reference_class.py`` runs the *reference's* ``WATS`` class
  (calibration/WATS.py:76-170) with only ``graph_wavelet_features`` swapped,
* ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's stock
  ``graph_wavelet_features`` (calibration/WATS.py:39-74).

Run by ``__graft_entry__.build()`` when ``/root/reference`` is mounted.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("EGNN_REFERENCE_ROOT", "/root/reference")

# module path relative to the reference root -> why the path needs it
MODULES = {
    "calibration/WATS.py": "the hot path and the calibrator (WATS.py:24-170)",
    "calibration/utils.py": "accuracy() used by calib_train (calibration/utils.py:139-167)",
    "utils/ece.py": "class-wise ECE of the downstream comparison (utils/ece.py:8-89)",
    "src/gnn/model.py": "CompatibleGCN base model (src/gnn/model.py:7-53)",
}


def build(verbose: bool = True) -> bool:
    """Compile the modules; returns False (and does nothing) when the reference is not mounted."""
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "calibration", "WATS.py")):
        if verbose:
            print(f"oracle/_ref: {REFERENCE_ROOT} not mounted, keeping whatever is already staged")
        return False
    for rel in MODULES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(OUT, os.path.splitext(rel)[0] + ".bin")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile keeps the reference path in tracebacks; unchecked-hash pyc needs no source beside it
        py_compile.compile(src, cfile=dst, dfile=src, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "STAGED"), "w") as fh:
        fh.write("byte-compiled by oracle/build_ref.py with python %d.%d from %s\n" %
                 (sys.version_info[0], sys.version_info[1], REFERENCE_ROOT))
        for rel in MODULES:
            fh.write(rel + "\n")
    if verbose:
        print(f"oracle/_ref: {len(MODULES)} reference modules byte-compiled into {OUT}")
    return True


if __name__ == "__main__":
    build()
