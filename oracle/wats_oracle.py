"""CPU oracle for the Chebyshev graph-wavelet feature path of WATS.

TEST INFRASTRUCTURE ONLY.  Nothing under ``efficient-gnn_b200/`` may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the timed CPU arm, never as the product path.

What it restates (citations relative to the reference repo):

* ``calibration/WATS.py:24-27``  compute_normalized_laplacian -> scipy
  ``csgraph.laplacian(adj, normed=True)``.  scipy is a third-party dependency
  of the reference (``requirements.txt:12``: ``scipy>=1.10.0``, unpinned; the
  build image carries 1.18.1).  Its ``_laplacian_sparse`` algorithm is restated
  in :func:`normalized_laplacian_parts` from the published semantics: in-degree
  (axis=0) minus the diagonal, isolated nodes get weight 1, two successive
  float32 divisions per stored entry, diagonal overwritten by ``1 - isolated``.
* ``calibration/WATS.py:55``     rescale ``(2/lambda_max) L - I`` (float64).
* ``calibration/WATS.py:58-59``  ``X0 = log1p(rowsum(A))`` (float32, self loops
  counted).
* ``calibration/WATS.py:29-37``  three-term Chebyshev recurrence, float64.
* ``calibration/WATS.py:65-68``  ``alpha_i = exp(-s i)``; ``S = sum alpha_i T_i``.
* ``calibration/WATS.py:71-72``  row L1 normalisation ``S / (|S|_1 + 1e-8)``.
* ``utils/ece.py:8-89``          class-wise ECE (bins with < 4 samples skipped).
* ``benchmark_calibration_methods.py:100-127`` accuracy / confidence / ECE.

Pinning: the reference has no tests or golden vectors (parity is unpinned by
the reference itself).  This oracle is pinned by ``oracle/make_golden.py``,
which imports the unmodified reference from ``/root/reference`` inside the
build container, runs it on seeded synthetic graphs and stores its outputs in
``tests/golden/*.npz``; ``tests/test_oracle.py`` then checks this restatement
against those files (bit-exact for the Laplacian entries and X0, <= 1e-13 for
the float64 stages), and against the live reference when it is mounted.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

__all__ = [
    "as_csr32",
    "normalized_laplacian_parts",
    "rescaled_laplacian",
    "input_signal",
    "chebyshev_orders",
    "heat_coefficients",
    "wavelet_parts",
    "wavelet_features",
    "classwise_ece",
    "average_ece",
    "evaluate_probs",
]


def as_csr32(adj) -> sp.csr_matrix:
    """Canonical CSR/float32 view of whatever adjacency the caller holds.

    Mirrors the only call site of the reference, ``csr_matrix(adj.cpu().numpy())``
    (calibration/WATS.py:99): explicit zeros are not stored, duplicates summed.
    """
    if sp.issparse(adj):
        m = sp.csr_matrix(adj, dtype=np.float32, copy=True)
    else:
        m = sp.csr_matrix(np.asarray(adj, dtype=np.float32))
    m.sum_duplicates()
    m.eliminate_zeros()
    m.sort_indices()
    return m


def normalized_laplacian_parts(adj: sp.csr_matrix):
    """scipy ``csgraph.laplacian(adj, normed=True)`` restated on raw CSR arrays.

    Returns ``(rows, cols, vals32, w32, isolated)``: the off-diagonal entries
    ``-a_ij / w_i / w_j`` (float32, two successive divisions, stored self-loops
    dropped) plus the weight vector ``w = sqrt(colsum - diag)`` (1 where that
    is 0) and the isolated mask.  The diagonal of L is ``1 - isolated``.
    Follows scipy/sparse/csgraph/_laplacian.py::_laplacian_sparse as called by
    calibration/WATS.py:26.
    """
    a = as_csr32(adj)
    n = a.shape[0]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(a.indptr))
    cols = a.indices.astype(np.int64)
    data = a.data.astype(np.float32)
    # in-degree (axis=0) accumulated in float32 like scipy's ones @ A.
    colsum = np.zeros(n, dtype=np.float32)
    np.add.at(colsum, cols, data)
    diag = np.zeros(n, dtype=np.float32)
    on_diag = rows == cols
    diag[rows[on_diag]] = data[on_diag]
    w = (colsum - diag).astype(np.float32)
    isolated = w == 0
    w = np.where(isolated, np.float32(1), np.sqrt(w)).astype(np.float32)
    off = ~on_diag
    r, c, v = rows[off], cols[off], data[off].copy()
    v /= w[r]
    v /= w[c]
    v *= np.float32(-1)
    return r, c, v, w, isolated


def rescaled_laplacian(adj, lambda_max: float = 2.0) -> sp.csr_matrix:
    """``(2/lambda_max) * L - I`` as float64 CSR (calibration/WATS.py:55).

    The float32 Laplacian entries are promoted to float64 *after* rounding, as
    scipy does when the float32 COO meets the float64 identity.
    """
    r, c, v, _w, isolated = normalized_laplacian_parts(adj)
    n = isolated.shape[0]
    scale = 2.0 / float(lambda_max)
    diag_l = (1 - isolated).astype(np.float32).astype(np.float64)
    rr = np.concatenate([r, np.arange(n)])
    cc = np.concatenate([c, np.arange(n)])
    vv = np.concatenate([scale * v.astype(np.float64), scale * diag_l - 1.0])
    m = sp.coo_matrix((vv, (rr, cc)), shape=(n, n)).tocsr()
    return m


def input_signal(adj) -> np.ndarray:
    """``log1p(A.sum(axis=1))`` as an ``[N,1]`` float32 column (WATS.py:58-59)."""
    a = as_csr32(adj)
    deg = np.asarray(a.sum(axis=1)).ravel()
    return np.log1p(deg).reshape(-1, 1)


def chebyshev_orders(lt, k: int, x0: np.ndarray):
    """``[T_0 .. T_k]`` with ``T_1 = L~ T_0``, ``T_i = 2 L~ T_{i-1} - T_{i-2}``
    (calibration/WATS.py:29-37).  ``lt`` is float64 CSR so every order >= 1 is
    float64 regardless of the dtype of ``x0``."""
    out = [x0]
    if k > 0:
        out.append(lt @ x0)
    for _ in range(2, k + 1):
        out.append(2 * lt @ out[-1] - out[-2])
    return out


def heat_coefficients(k: int, s) -> np.ndarray:
    """``alpha[j, i] = exp(-s_j * i)`` for i = 0..k (calibration/WATS.py:65)."""
    s = np.atleast_1d(np.asarray(s, dtype=np.float64))
    return np.exp(-s[:, None] * np.arange(k + 1, dtype=np.float64)[None, :])


def wavelet_parts(adj, k: int = 3, s=0.8, x0=None, lambda_max: float = 2.0):
    """Every intermediate of the path: ``dict(X0, T=[...], S=[S_j], H=[H_j])``.

    ``s`` may be a scalar (reference behaviour) or a sequence of scales; every
    scale is combined and normalised exactly as the reference does for one.
    """
    lt = rescaled_laplacian(adj, lambda_max)
    if x0 is None:
        x0 = input_signal(adj)
    else:
        x0 = np.asarray(x0)
        if x0.ndim == 1:
            x0 = x0.reshape(-1, 1)
    orders = chebyshev_orders(lt, k, x0)
    alpha = heat_coefficients(k, s)
    s_list, h_list = [], []
    for a in alpha:
        comb = sum(a[i] * orders[i] for i in range(k + 1))
        comb = np.asarray(comb, dtype=np.float64)
        norm = np.abs(comb).sum(axis=1, keepdims=True) + 1e-8
        s_list.append(comb)
        h_list.append(comb / norm)
    return {"X0": x0, "T": orders, "S": s_list, "H": h_list, "alpha": alpha}


def wavelet_features(adj, k: int = 3, s=0.8, x0=None, lambda_max: float = 2.0):
    """Drop-in restatement of ``graph_wavelet_features`` (WATS.py:39-74).

    Scalar ``s`` -> ``[N,F]`` float64; sequence -> ``[N, len(s)*F]``.
    """
    parts = wavelet_parts(adj, k, s, x0, lambda_max)
    if np.ndim(s) == 0:
        return parts["H"][0]
    return np.concatenate(parts["H"], axis=1)


# --------------------------------------------------------------------------- #
# downstream metrics (oracle side of the accuracy / ECE / confidence parity)  #
# --------------------------------------------------------------------------- #
def classwise_ece(probs: np.ndarray, labels: np.ndarray, pos_class: int,
                  n_bins: int = 10) -> float:
    """One-vs-rest ECE of ``probs[:, pos_class]`` (utils/ece.py:8-62):
    right-closed bins, bins holding fewer than 4 samples are skipped."""
    p = probs[:, pos_class]
    hit = labels == pos_class
    edges = np.linspace(0, 1, n_bins + 1)
    which = np.digitize(p, edges, right=True) - 1
    total = 0.0
    for b in range(n_bins):
        m = which == b
        if m.sum() < 4:
            continue
        total += abs(p[m].mean() - hit[m].mean()) * m.mean()
    return float(total)


def average_ece(probs, labels, n_classes, n_bins: int = 10) -> float:
    """Mean of the class-wise ECEs (utils/ece.py:64-89, ``logits=False``)."""
    return float(np.mean([classwise_ece(probs, labels, c, n_bins)
                          for c in range(n_classes)]))


def evaluate_probs(log_probs: np.ndarray, labels: np.ndarray, mask: np.ndarray):
    """(accuracy, mean max-probability, class-wise ECE) on ``mask``
    (benchmark_calibration_methods.py:100-127)."""
    probs = np.exp(log_probs)[mask]
    y = labels[mask]
    acc = float((probs.argmax(axis=1) == y).mean())
    conf = float(probs.max(axis=1).mean())
    ece = average_ece(probs, y, probs.shape[1])
    return acc, conf, ece
