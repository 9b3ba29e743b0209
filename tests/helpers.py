"""Shared test helpers: golden-file loading and the parity norms of SURVEY 8c."""
import os

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

FEATURE_CASES = ["kat_path", "kat_path_loops", "cora_noloop", "cora_loops", "pubmed_noloop",
                 "cora_k0", "cora_k1", "cora_k6_s04", "directed_weighted", "cora_wide8",
                 "small_wide130"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    n = int(z["n"])
    adj = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
    k = int(z["k"])
    return dict(name=name, n=n, k=k, s=float(z["s"]), adj=adj,
                lt=sp.csr_matrix((z["lt_data"], z["lt_indices"], z["lt_indptr"]), shape=(n, n)),
                X0=z["X0"], T=[z[f"T{i}"] for i in range(k + 1)], S=z["S"], H=z["H"],
                custom_x0=bool(z["custom_x0"]))


def rel_max_err(got, ref):
    """||got-ref||_inf / ||ref||_inf (the per-order parity norm)."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    denom = np.abs(ref).max()
    if denom == 0:
        return float(np.abs(got).max())
    return float(np.abs(got - ref).max() / denom)


ELEMENT_FLOOR = 5e-2


def elementwise_ratio(got, ref, tol=1e-5, floor=None):
    """max over elements of |delta| / (tol * max(|ref|, ELEMENT_FLOOR * ||ref||_inf)); <= 1 passes.

    SURVEY 8c proposed a floor of 1e-3 * ||ref||_inf, i.e. an absolute error of
    1e-8 * ||ref||_inf - below float32 resolution (eps = 6e-8) wherever a row
    sum cancels, so no float32 implementation can meet it.  With the floor at
    5e-2 the bound is ~8 eps * ||ref||_inf absolute, which leaves room for the
    O(k^2 eps) growth of the three-term recurrence.  The norm-wise 1e-5 bound
    (north_star) is checked separately and is met with ~100x margin.
    ``floor`` overrides the floor (tests print the ratio at SURVEY's 1e-3 next to
    the enforced one so the gap stays visible; DESIGN.md section 2 records both)."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    floor = (ELEMENT_FLOOR if floor is None else floor) * np.abs(ref).max()
    return float((np.abs(got - ref) / (tol * np.maximum(np.abs(ref), floor) + 1e-300)).max())

