"""The CPU oracle against the reference's own outputs (tests/golden, generated
by oracle/make_golden.py from the unmodified reference) and, when mounted,
against the live reference.  No GPU."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import wats_oracle as orc
from oracle import ref_shim
from helpers import FEATURE_CASES, load_case


@pytest.mark.parametrize("name", FEATURE_CASES)
def test_laplacian_entries_bit_exact(name):
    c = load_case(name)
    lt = orc.rescaled_laplacian(c["adj"])
    lt.sort_indices()
    ref = c["lt"].copy()
    # the reference keeps explicit zeros on the diagonal of non-isolated nodes
    for m in (lt, ref):
        m.eliminate_zeros()
        m.sort_indices()
    assert np.array_equal(lt.indptr, ref.indptr)
    assert np.array_equal(lt.indices, ref.indices)
    assert np.array_equal(lt.data, ref.data)          # float32-rounded entries, exact


@pytest.mark.parametrize("name", FEATURE_CASES)
def test_orders_and_features(name):
    c = load_case(name)
    x0 = c["X0"] if c["custom_x0"] else None
    p = orc.wavelet_parts(c["adj"], k=c["k"], s=c["s"], x0=x0)
    assert p["X0"].dtype == c["X0"].dtype
    assert np.array_equal(p["X0"], c["X0"])
    for got, ref in zip(p["T"], c["T"]):
        assert got.dtype == ref.dtype
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13)
    np.testing.assert_allclose(p["S"][0], c["S"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(p["H"][0], c["H"], rtol=0, atol=1e-13)
    h = orc.wavelet_features(c["adj"], k=c["k"], s=c["s"], x0=x0)
    np.testing.assert_allclose(h, c["H"], rtol=0, atol=1e-13)


def test_known_answer_path_graph():
    """SURVEY 8c hand-checkable vector: path 0-1-2 plus isolated node 3."""
    a = np.zeros((4, 4), np.float32)
    a[0, 1] = a[1, 0] = a[1, 2] = a[2, 1] = 1
    p = orc.wavelet_parts(sp.csr_matrix(a))
    r = np.float32(1) / np.sqrt(np.float32(2))
    lt = orc.rescaled_laplacian(sp.csr_matrix(a)).toarray()
    expect = np.array([[0, -r, 0, 0], [-r, 0, -r, 0], [0, -r, 0, 0], [0, 0, 0, -1]], np.float64)
    np.testing.assert_allclose(lt, expect, atol=1e-7)
    np.testing.assert_allclose(p["X0"].ravel(), [0.6931472, 1.0986123, 0.6931472, 0], atol=1e-7)
    np.testing.assert_allclose(p["T"][1].ravel(), [-0.7768362, -0.98025813, -0.7768362, 0], atol=1e-7)
    np.testing.assert_allclose(p["T"][2].ravel(), [0.69314714, 1.09861223, 0.69314714, 0], atol=1e-7)
    np.testing.assert_allclose(p["T"][3].ravel(), [-0.77683609, -0.980258, -0.77683609, 0], atol=1e-7)
    np.testing.assert_allclose(p["H"][0].ravel(), [1, 1, 1, 0], atol=1e-7)


def test_multi_scale_is_per_scale_reference():
    c = load_case("cora_noloop")
    scales = [0.4, 0.8, 1.6]
    multi = orc.wavelet_features(c["adj"], k=3, s=scales)
    assert multi.shape == (c["n"], 3)
    for j, s in enumerate(scales):
        single = orc.wavelet_features(c["adj"], k=3, s=s)
        np.testing.assert_array_equal(multi[:, j:j + 1], single)
    np.testing.assert_allclose(multi[:, 1:2], c["H"], atol=1e-13)


def test_lambda_max_rescale():
    c = load_case("cora_loops")
    for lam in (1.5, 2.0, 2.5):
        lt = orc.rescaled_laplacian(c["adj"], lam)
        r, cc, v, w, iso = orc.normalized_laplacian_parts(c["adj"])
        lap = sp.coo_matrix((v.astype(np.float64), (r, cc)), shape=lt.shape).tocsr() \
            + sp.diags((1 - iso).astype(np.float64))
        want = (2.0 / lam) * lap - sp.identity(lt.shape[0])
        assert abs(lt - want).max() < 1e-15


def test_ece_matches_reference_formula():
    rng = np.random.default_rng(0)
    logits = rng.standard_normal((700, 5)) * 2
    probs = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    y = rng.integers(0, 5, 700)
    e = orc.average_ece(probs, y, 5)
    assert 0 < e < 1
    # one class, by hand
    p = probs[:, 2]
    hit = y == 2
    total = 0.0
    edges = np.linspace(0, 1, 11)
    for b in range(10):
        m = (p > edges[b]) & (p <= edges[b + 1])
        if m.sum() >= 4:
            total += abs(p[m].mean() - hit[m].mean()) * m.mean()
    assert abs(orc.classwise_ece(probs, y, 2) - total) < 1e-15


@pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference neither mounted nor staged in oracle/_ref")
@pytest.mark.parametrize("where", ["mounted", "staged"])
def test_against_live_reference(where):
    """The restatement against the reference's own code: the checkout in the build
    container, and the byte-compiled copy in oracle/_ref that travels to the GPU box."""
    root = ref_shim.REFERENCE_ROOT if where == "mounted" else ref_shim.STAGED_ROOT
    if (where == "mounted" and not ref_shim.reference_available()) or \
            (where == "staged" and not ref_shim.staged_available()):
        pytest.skip(f"reference not {where}")
    wats, _model, ece = ref_shim.load_reference(root)
    assert wats.__file__.startswith(root)
    from efficient_gnn_b200 import synth
    rp, ci, n = synth.synth_csr(synth.GraphShape("t", 1500, 9000, 3, 77, 1), self_loops=True)
    adj = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n))
    for k, s in ((3, 0.8), (5, 0.3)):
        ref = wats.graph_wavelet_features(adj, k=k, s=s)
        got = orc.wavelet_features(adj, k=k, s=s)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13)
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((400, 3))
    probs = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    y = rng.integers(0, 3, 400)
    assert abs(ece.calculate_average_ece(probs, y, 3, logits=False) - orc.average_ece(probs, y, 3)) < 1e-15
