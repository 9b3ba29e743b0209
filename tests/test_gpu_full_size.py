"""Parity against the oracle AT THE SIZES THE PUBLISHED NUMBERS ARE QUOTED ON
(BASELINE.json configs 3-5, bench.py's default workload and its wide
sub-records): the kernels that bench.py times, on the graphs it times them on.

* Reddit shape (232,965 nodes, 114.8 M stored entries, seed 5, self loops),
  F = 1, K = 3 through the SELL plan - forced, and as the default second-call
  path - against ``orc.chebyshev_orders`` on the full graph: T_1..T_3 and S.
* Reddit shape F = 64 and Physics shape F = 8415 through ``cheb_wide_kernel``:
  the GPU runs the full width, the oracle a spread of columns (the operator
  acts on every column independently, so a column subset is an exact check of
  those columns; the full-width oracle would need minutes of scipy time).
* arxiv shape F = 128 through ``cheb_wide_kernel`` against the oracle at full
  width.

Tolerances are SURVEY 8c's: per order ||gpu - ref||_inf / ||ref||_inf <= 1e-5
plus the element-wise bound of tests/helpers.py; the ratio at SURVEY's original
1e-3 floor is printed beside it (DESIGN.md section 2 records the worst cases).
"""
import time

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
from oracle import wats_oracle as orc
from helpers import elementwise_ratio, rel_max_err

pytestmark = pytest.mark.gpu

TOL = 1e-5
K = 3
SCALES = [0.8, 1.6]


def host_csr(g):
    return sp.csr_matrix((np.ones(g.nnz, np.float32), g.colidx.cpu().numpy(), g.rowptr.cpu().numpy()),
                         shape=(g.n, g.n))


def check_orders(got_orders, ref_orders, tag):
    """Norm-wise and element-wise bound per order; returns the printed summary."""
    rows = []
    for i, (got, ref) in enumerate(zip(got_orders, ref_orders)):
        g = got.cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
        err = rel_max_err(g, ref)
        ratio = elementwise_ratio(g, ref, TOL)
        at_survey_floor = elementwise_ratio(g, ref, TOL, floor=1e-3)
        rows.append((i, err, ratio, at_survey_floor))
        assert err <= TOL, f"{tag} order {i}: rel max err {err:.3e}"
        assert ratio <= 1.0, f"{tag} order {i}: element-wise bound exceeded x{ratio:.2f}"
    print(f"\n[{tag}] " + "; ".join(f"T{i}: rel {e:.2e}, elem x{r:.2f} (x{r3:.1f} at the 1e-3 floor)"
                                    for i, e, r, r3 in rows))
    return rows


def combine(orders, s):
    alpha = orc.heat_coefficients(len(orders) - 1, s)
    return [sum(a[i] * np.asarray(orders[i], dtype=np.float64) for i in range(len(orders))) for a in alpha]


@pytest.fixture(scope="module")
def reddit():
    """The graph bench.py times by default, its host copy and the oracle's rescaled Laplacian
    (float64 CSR, built once: ~115 M entries)."""
    rp, ci, n = synth.synth_csr("reddit", self_loops=True, device="cuda")
    g = egnn.CsrGraph(rp, ci, None, n)
    adj = host_csr(g)
    t0 = time.perf_counter()
    lt = orc.rescaled_laplacian(adj)
    x0 = orc.input_signal(adj)
    print(f"\noracle operator at Reddit size: {time.perf_counter() - t0:.1f} s")
    return g, adj, lt, x0


def test_reddit_f1_sell_plan_vs_oracle(reddit):
    g, adj, lt, x0 = reddit
    ref_T = orc.chebyshev_orders(lt, K, x0)
    ref_S = combine(ref_T, SCALES)
    # forced plan
    res = egnn.graph_wavelet_features(g, k=K, s=SCALES, return_parts=True, _use_sell=True)
    assert g.has_sell_plan()
    check_orders(res.orders, ref_T, "reddit F=1 SELL")
    check_orders([res.combined[:, j, :] for j in range(len(SCALES))], ref_S, "reddit F=1 SELL S")
    # degree signal and normaliser of the benchmarked graph
    np.testing.assert_allclose(g.x0.cpu().numpy(), x0.ravel(), rtol=2e-7, atol=0)     # log1pf within 2 ulp
    # the generic kernel on the same graph (what a one-shot call runs)
    gen = egnn.graph_wavelet_features(g, k=K, s=SCALES, return_parts=True, _use_sell=False)
    check_orders(gen.orders, ref_T, "reddit F=1 generic")
    # final features: sign(S) wherever S is not within rounding of zero
    feats = egnn.graph_wavelet_features(g, k=K, s=0.8, _use_sell=True).cpu().numpy()
    s_ref = ref_S[0]
    sure = np.abs(s_ref) > 1e-4 * np.abs(s_ref).max()
    h_ref = (s_ref / (np.abs(s_ref) + 1e-8)).astype(np.float32)
    np.testing.assert_allclose(feats[sure], h_ref[sure], rtol=0, atol=1e-6)


def test_reddit_default_second_call_path_vs_oracle():
    """No private switches: the first call of a fresh graph runs the generic kernel, the
    second builds the plan and runs the SELL kernels (what the UGCA loop and bench.py get)."""
    rp, ci, n = synth.synth_csr("reddit", self_loops=True, device="cuda")
    g = egnn.CsrGraph(rp, ci, None, n)
    adj = host_csr(g)
    ref_T = orc.chebyshev_orders(orc.rescaled_laplacian(adj), K, orc.input_signal(adj))
    first = egnn.graph_wavelet_features(g, k=K, s=0.8, return_parts=True)
    assert not g.has_sell_plan()
    second = egnn.graph_wavelet_features(g, k=K, s=0.8, return_parts=True)
    assert g.has_sell_plan()
    check_orders(first.orders, ref_T, "reddit default call 1")
    check_orders(second.orders, ref_T, "reddit default call 2")


def test_reddit_f64_wide_kernel_vs_oracle_columns(reddit):
    g, adj, lt, _ = reddit
    f = 64
    x = torch.randn(g.n, f, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    res = egnn.graph_wavelet_features(g, k=K, s=SCALES, X0=x, return_parts=True)
    cols = [0, 21, 42, 63]
    ref_T = orc.chebyshev_orders(lt, K, x[:, cols].cpu().numpy())
    check_orders([t[:, cols] for t in res.orders], ref_T, "reddit F=64 wide")
    check_orders([res.combined[:, j, :][:, cols] for j in range(len(SCALES))], combine(ref_T, SCALES),
                 "reddit F=64 wide S")
    # the fused normalising pass against the parts pass
    fused = egnn.graph_wavelet_features(g, k=K, s=SCALES, X0=x)
    want = res.combined / (res.combined.abs().sum(dim=2, keepdim=True) + 1e-8)
    assert torch.allclose(fused, want.reshape(g.n, -1), rtol=0, atol=2e-6)


def test_arxiv_f128_wide_kernel_vs_oracle():
    rp, ci, n = synth.synth_csr("arxiv", self_loops=True, device="cuda")
    g = egnn.CsrGraph(rp, ci, None, n)
    adj = host_csr(g)
    x = torch.randn(n, 128, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    res = egnn.graph_wavelet_features(g, k=K, s=SCALES, X0=x, return_parts=True)
    p = orc.wavelet_parts(adj, k=K, s=SCALES, x0=x.cpu().numpy())
    check_orders(res.orders, p["T"], "arxiv F=128 wide")
    check_orders([res.combined[:, j, :] for j in range(len(SCALES))], p["S"], "arxiv F=128 wide S")
    h = res.features.cpu().numpy()
    for j, rh in enumerate(p["H"]):
        np.testing.assert_allclose(h[:, j * 128:(j + 1) * 128], rh.astype(np.float32), rtol=0, atol=2e-5)


def test_physics_f8415_wide_kernel_vs_oracle_columns():
    sh = synth.SHAPES["physics"]
    rp, ci, n = synth.synth_csr("physics", self_loops=True, device="cuda")
    g = egnn.CsrGraph(rp, ci, None, n)
    adj = host_csr(g)
    f = sh.f_wide
    assert f == 8415 and f % 4 == 3                 # the unaligned-row instantiation
    x = torch.randn(n, f, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    res = egnn.graph_wavelet_features(g, k=K, s=0.8, X0=x, return_parts=True)
    # first / last columns of the row, both sides of feature-tile boundaries, the ragged tail
    cols = [0, 1, 127, 128, 129, 4095, 4096, 4207, 6399, 6400, 8191, 8192, 8411, 8412, 8413, 8414]
    ref_T = orc.chebyshev_orders(orc.rescaled_laplacian(adj), K, x[:, cols].cpu().numpy())
    check_orders([t[:, cols] for t in res.orders], ref_T, "physics F=8415 wide")
    check_orders([res.combined[:, 0, :][:, cols]], combine(ref_T, 0.8), "physics F=8415 wide S")
    fused = egnn.graph_wavelet_features(g, k=K, s=0.8, X0=x)
    want = res.combined[:, 0, :] / (res.combined[:, 0, :].abs().sum(dim=1, keepdim=True) + 1e-8)
    assert torch.allclose(fused, want, rtol=0, atol=2e-6)
