"""Seeded synthetic graphs of the five shapes BASELINE.json names.

Datasets are not available offline, so every measurement and parity run uses a
Chung-Lu style power-law graph with the node / edge counts of the named
dataset.  Two adjacency conventions reach the hot path
in the reference and both are generated:

* ``self_loops=False`` - benchmark-script convention
  (calibration/utils.py:19-25 via benchmark_calibration_methods.py:53);
* ``self_loops=True``  - attack-script convention, symmetrised with unit
  diagonal (ugca_calib_attack.py:42-47).

The generator is written in torch ops so the Reddit shape (114.6 M stored
entries) can be produced directly in HBM; the random stream depends on the
device type, so parity tests always generate on the CPU and copy.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

__all__ = ["SHAPES", "GraphShape", "synth_edges", "synth_csr", "synth_labels"]


@dataclass(frozen=True)
class GraphShape:
    name: str
    n: int            # nodes
    nnz: int          # stored off-diagonal entries (both directions)
    n_classes: int
    seed: int         # rng seed = config index 1..5
    f_wide: int       # native feature width used for the wide-F runs


SHAPES = {
    "cora": GraphShape("cora", 2_708, 10_556, 7, 1, 1),
    "pubmed": GraphShape("pubmed", 19_717, 88_648, 3, 2, 1),
    "physics": GraphShape("physics", 34_493, 495_924, 5, 3, 8_415),
    "arxiv": GraphShape("arxiv", 169_343, 2_332_486, 40, 4, 128),
    "reddit": GraphShape("reddit", 232_965, 114_615_892, 41, 5, 64),
}


def _node_weights(n: int, skew: float, hub_ratio: float, device) -> torch.Tensor:
    """w_i ~ (i + i0)^-skew, i0 chosen so max/mean expected degree ~ hub_ratio."""
    i = torch.arange(n, dtype=torch.float64, device=device)
    lo, hi = 1e-3, float(n)
    for _ in range(60):           # bisection on the offset
        mid = (lo * hi) ** 0.5
        w = (i + mid) ** (-skew)
        if (w[0] / w.mean()).item() > hub_ratio:
            lo = mid
        else:
            hi = mid
    return (i + hi) ** (-skew)


def synth_edges(n: int, nnz: int, seed: int, *, device="cpu", skew: float = 0.6,
                hub_ratio: float = 45.0) -> torch.Tensor:
    """Undirected simple graph as a sorted ``[2, ~nnz]`` int64 edge list (both
    directions stored, no self edges, no duplicates)."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    want = nnz // 2
    w = _node_weights(n, skew, hub_ratio, device)
    perm = torch.randperm(n, generator=gen, device=device)
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1]
    keys = torch.empty(0, dtype=torch.int64, device=device)
    draw = int(want * 1.15) + 16
    for _ in range(8):
        u = torch.rand(draw, generator=gen, device=device, dtype=torch.float64)
        v = torch.rand(draw, generator=gen, device=device, dtype=torch.float64)
        a = perm[torch.searchsorted(cdf, u).clamp_(max=n - 1)]
        b = perm[torch.searchsorted(cdf, v).clamp_(max=n - 1)]
        del u, v
        lo_, hi_ = torch.minimum(a, b), torch.maximum(a, b)
        ok = lo_ != hi_
        keys = torch.unique(torch.cat([keys, (lo_ * n + hi_)[ok]]))
        del a, b, lo_, hi_, ok
        if keys.numel() >= want:
            break
        draw = int((want - keys.numel()) * 1.3) + 16
    if keys.numel() > want:       # drop a seeded random subset down to the target
        keep = torch.randperm(keys.numel(), generator=gen, device=device)[:want]
        keys = keys[keep.sort().values]
    lo_, hi_ = keys // n, keys % n
    both = torch.cat([lo_ * n + hi_, hi_ * n + lo_]).sort().values
    return torch.stack([both // n, both % n])


def synth_csr(shape, *, self_loops: bool = False, device="cpu", scale: float = 1.0):
    """CSR arrays ``(rowptr int32 [N+1], colidx int32 [nnz], N)`` of a named
    shape (or a :class:`GraphShape`); ``scale`` < 1 shrinks nodes and entries
    together (same mean degree) for bounded CPU-baseline samples."""
    sh = SHAPES[shape] if isinstance(shape, str) else shape
    n = max(8, int(round(sh.n * scale)))
    nnz = max(8, int(round(sh.nnz * scale)))
    e = synth_edges(n, nnz, sh.seed, device=device)
    if self_loops:
        d = torch.arange(n, dtype=torch.int64, device=e.device)
        key = torch.cat([e[0] * n + e[1], d * n + d]).sort().values
        e = torch.stack([key // n, key % n])
    counts = torch.bincount(e[0], minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=e.device)
    rowptr[1:] = torch.cumsum(counts, 0)
    return rowptr.to(torch.int32), e[1].to(torch.int32), n


def synth_labels(n: int, n_classes: int, seed: int = 42):
    """Stub base-model outputs for the downstream parity runs:
    labels uniform over C, logits ``3*onehot(y) + N(0,1)``, 500 validation and
    1000 test nodes (Planetoid-like), all seeded on the CPU."""
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, n_classes, (n,), generator=g)
    logits = 3.0 * torch.nn.functional.one_hot(y, n_classes).float() \
        + torch.randn(n, n_classes, generator=g)
    order = torch.randperm(n, generator=g)
    n_val = min(500, n // 4)
    n_test = min(1000, n // 2)
    val = torch.zeros(n, dtype=torch.bool)
    test = torch.zeros(n, dtype=torch.bool)
    val[order[:n_val]] = True
    test[order[n_val:n_val + n_test]] = True
    return y, logits, val, test
