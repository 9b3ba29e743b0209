"""efficient-gnn-b200: B200-native Chebyshev graph-wavelet features for WATS.

Only the hot path of CaptainCuong/Efficient-GNN is implemented here
(calibration/WATS.py:24-130): CUDA kernels behind a C-ABI shared library
(``csrc/`` -> ``lib/libegnn_b200.so``, declared in ``include/egnn_b200.h``)
and the PyTorch host code that mirrors the reference's Python interface.
Importing the package does not need a GPU; calling into it does (there is no
CPU fallback).
"""
__version__ = "0.1.0"

from . import _cabi  # noqa: F401
from ._cabi import EgnnError  # noqa: F401
from .graph import CsrGraph, as_graph  # noqa: F401
from .metrics import calibration_metrics  # noqa: F401
from .surrogate import SparseGCNSurrogate, StructureGradient  # noqa: F401
from .wats import (  # noqa: F401
    WATS,
    LaplacianOperator,
    WaveletResult,
    WaveletSession,
    accuracy,
    chebyshev_polynomials,
    compute_normalized_laplacian,
    graph_wavelet_features,
    heat_coefficients,
)
