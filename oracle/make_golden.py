"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run inside the build container only (``/root/reference`` is mounted there):

    python oracle/make_golden.py

Every file stores the inputs (CSR of the adjacency) and the reference's own
outputs for the hot path: the rescaled Laplacian as built by
``compute_normalized_laplacian`` + the rescale line (calibration/WATS.py:53-55),
``X0`` (:58-59), ``chebyshev_polynomials`` orders (:29-37), the combination S
(:65-68) and ``graph_wavelet_features`` (:39-74).  The downstream files hold the
accuracy / confidence / ECE triple of the reference ``WATS`` class
(calibration/WATS.py:76-170) evaluated as benchmark_calibration_methods.py:100-127
does, with ``torch.manual_seed(42)``.

TEST INFRASTRUCTURE ONLY - the product never reads these files.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from efficient_gnn_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def csr_from_shape(name, self_loops):
    rp, ci, n = synth.synth_csr(name, self_loops=self_loops)
    data = np.ones(ci.numel(), dtype=np.float32)
    return sp.csr_matrix((data, ci.numpy(), rp.numpy()), shape=(n, n))


def reference_parts(wats, adj, k, s, x0=None):
    """Run the reference line by line so the intermediates can be recorded."""
    n = adj.shape[0]
    lap = wats.compute_normalized_laplacian(adj)                 # WATS.py:53
    lt = (2 / 2.0) * lap - sp.identity(n)                        # WATS.py:55
    if x0 is None:
        deg = np.array(adj.sum(axis=1)).flatten()                # WATS.py:58
        x0 = np.log1p(deg).reshape(-1, 1)                        # WATS.py:59
    orders = wats.chebyshev_polynomials(lt, k, x0)               # WATS.py:62
    alpha = [np.exp(-s * i) for i in range(k + 1)]               # WATS.py:65
    comb = sum(alpha[i] * orders[i] for i in range(k + 1))       # WATS.py:68
    lt = sp.csr_matrix(lt)
    lt.sort_indices()
    return lt, x0, orders, comb


def save_case(wats, name, adj, k=3, s=0.8, x0=None):
    adj = sp.csr_matrix(adj, dtype=np.float32)
    adj.sort_indices()
    lt, x0_used, orders, comb = reference_parts(wats, adj, k, s, x0)
    payload = dict(
        n=np.int64(adj.shape[0]), k=np.int64(k), s=np.float64(s),
        indptr=adj.indptr.astype(np.int64), indices=adj.indices.astype(np.int32),
        data=adj.data.astype(np.float32),
        lt_indptr=lt.indptr.astype(np.int64), lt_indices=lt.indices.astype(np.int32),
        lt_data=lt.data.astype(np.float64),
        X0=np.asarray(x0_used), S=np.asarray(comb, dtype=np.float64),
        custom_x0=np.bool_(x0 is not None),
    )
    for i, t in enumerate(orders):
        payload[f"T{i}"] = np.asarray(t)
    if x0 is None:
        payload["H"] = np.asarray(wats.graph_wavelet_features(adj, k=k, s=s))
    else:   # the reference driver takes no X0: normalise S as WATS.py:71-72 does
        payload["H"] = comb / (np.linalg.norm(comb, ord=1, axis=1, keepdims=True) + 1e-8)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **payload)
    print(f"{name}: N={adj.shape[0]} nnz={adj.nnz} k={k} s={s} F={x0_used.shape[1]}")


def dense_from_csr(adj):
    return torch.tensor(adj.toarray(), dtype=torch.float32)


class FixedLogits(torch.nn.Module):
    """Stub base model of SURVEY 8d: returns fixed logits, ignores x and adj."""

    def __init__(self, logits):
        super().__init__()
        self.register_buffer("logits", logits)

    def forward(self, x, adj):
        return self.logits


def evaluate(ece_mod, model, x, y, adj, test_mask):
    """benchmark_calibration_methods.py:100-127 (accuracy, confidence, ECE)."""
    model.eval()
    with torch.no_grad():
        probs = model(x, adj).exp()
        tp, tl = probs[test_mask], y[test_mask]
        acc = (tp.argmax(dim=1) == tl).float().mean().item()
        conf = tp.max(dim=1)[0].mean().item()
        ece = ece_mod.calculate_average_ece(tp.cpu().numpy(), tl.cpu().numpy(),
                                            tp.shape[1], logits=False)
    return acc, conf, float(ece)


def save_downstream(wats, model_mod, ece_mod, name, shape, self_loops, use_gcn):
    sh = synth.SHAPES[shape]
    adj_csr = csr_from_shape(shape, self_loops)
    adj = dense_from_csr(adj_csr)
    y, logits, val, test = synth.synth_labels(sh.n, sh.n_classes, seed=42)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(sh.n, 16, generator=g)
    torch.manual_seed(42)                                    # benchmark :166-167
    np.random.seed(42)
    if use_gcn:
        base = model_mod.CompatibleGCN(nfeat=16, nclass=sh.n_classes)
    else:
        base = FixedLogits(logits)
    with contextlib.redirect_stdout(io.StringIO()):
        cal = wats.WATS(base, x, y, adj, val)
    acc, conf, ece = evaluate(ece_mod, cal, x, y, adj, test)
    sd = {k: v.detach().cpu().numpy() for k, v in cal.net.state_dict().items()}
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        shape=np.str_(shape), self_loops=np.bool_(self_loops), use_gcn=np.bool_(use_gcn),
        acc=np.float64(acc), conf=np.float64(conf), ece=np.float64(ece),
        wavelet_feats=cal.wavelet_feats.cpu().numpy(),
        **{f"net.{k}": v for k, v in sd.items()})
    print(f"{name}: acc={acc:.4f} conf={conf:.4f} ece={ece:.4f}")


def main():
    os.makedirs(OUT, exist_ok=True)
    wats, model_mod, ece_mod = load_reference()

    # hand-checkable known-answer graph of SURVEY 8c: path 0-1-2 plus isolated 3
    path = np.zeros((4, 4), dtype=np.float32)
    path[0, 1] = path[1, 0] = path[1, 2] = path[2, 1] = 1
    save_case(wats, "kat_path", sp.csr_matrix(path))
    save_case(wats, "kat_path_loops", sp.csr_matrix(path + np.eye(4, dtype=np.float32)))

    save_case(wats, "cora_noloop", csr_from_shape("cora", False))
    save_case(wats, "cora_loops", csr_from_shape("cora", True))
    save_case(wats, "pubmed_noloop", csr_from_shape("pubmed", False))
    save_case(wats, "cora_k0", csr_from_shape("cora", False), k=0)
    save_case(wats, "cora_k1", csr_from_shape("cora", False), k=1)
    save_case(wats, "cora_k6_s04", csr_from_shape("cora", True), k=6, s=0.4)

    # directed + weighted + partial self-loops: in-degree normalisation, a
    # node with out-edges only, an isolated node
    rng = np.random.default_rng(11)
    n = 300
    dense = (rng.random((n, n)) < 0.03) * rng.uniform(0.25, 3.0, (n, n))
    dense[np.arange(0, n, 7), np.arange(0, n, 7)] = 1.5
    dense[:, 5] = 0.0          # node 5: no in-edges (w=0 -> isolated) but out-edges
    dense[9, :] = 0.0
    dense[:, 9] = 0.0          # node 9: fully isolated
    save_case(wats, "directed_weighted", sp.csr_matrix(dense.astype(np.float32)))

    # wide input signal through the reference recurrence (F = 8 and F = 130)
    adj = csr_from_shape("cora", False)
    rng = np.random.default_rng(3)
    save_case(wats, "cora_wide8", adj,
              x0=rng.standard_normal((adj.shape[0], 8)).astype(np.float32))
    rp, ci, n = synth.synth_csr(synth.GraphShape("small", 256, 2400, 4, 9, 130))
    small = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n))
    save_case(wats, "small_wide130", small, k=4, s=1.6,
              x0=rng.standard_normal((n, 130)).astype(np.float32))

    save_downstream(wats, model_mod, ece_mod, "downstream_cora_stub", "cora", False, False)
    save_downstream(wats, model_mod, ece_mod, "downstream_cora_gcn", "cora", True, True)


if __name__ == "__main__":
    main()
