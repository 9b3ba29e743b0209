"""efficient-gnn-b200: B200-native Chebyshev graph-wavelet features for WATS.

Only the hot path of CaptainCuong/Efficient-GNN is implemented here
(calibration/WATS.py:24-130): CUDA kernels behind a C-ABI shared library
(``csrc/`` -> ``lib/libegnn_b200.so``, declared in ``include/egnn_b200.h``)
and the PyTorch host code that mirrors the reference's Python interface.
"""
__version__ = "0.1.0"
