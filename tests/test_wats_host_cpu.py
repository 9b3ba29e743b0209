"""The WATS module mirror (efficient-gnn_b200/wats.py) against the reference
class on the CPU: same seed, features injected from the oracle, so the only
thing under test is the host-side mirror (net layout, temperature head,
training loop, state_dict keys).  Golden numbers come from the unmodified
reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
from oracle import wats_oracle as orc
from helpers import GOLDEN
from models_for_tests import DenseGCN, FixedLogits


def _setup(shape, self_loops):
    sh = synth.SHAPES[shape]
    rp, ci, n = synth.synth_csr(shape, self_loops=self_loops)
    adj_csr = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n))
    adj = torch.tensor(adj_csr.toarray(), dtype=torch.float32)
    y, logits, val, test = synth.synth_labels(sh.n, sh.n_classes, seed=42)
    x = torch.randn(sh.n, 16, generator=torch.Generator().manual_seed(7))
    return sh, adj_csr, adj, x, y, logits, val, test


@pytest.mark.skipif(torch.cuda.is_available(), reason="golden numbers were produced on the CPU")
@pytest.mark.parametrize("name,use_gcn,self_loops", [("downstream_cora_stub", False, False),
                                                     ("downstream_cora_gcn", True, True)])
def test_wats_mirror_reproduces_reference_numbers(name, use_gcn, self_loops):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    sh, adj_csr, adj, x, y, logits, val, test = _setup("cora", self_loops)
    feats = orc.wavelet_features(adj_csr).astype(np.float32)
    np.testing.assert_array_equal(feats, g["wavelet_feats"])
    torch.manual_seed(42)
    np.random.seed(42)
    base = DenseGCN(16, sh.n_classes) if use_gcn else FixedLogits(logits)
    cal = egnn.WATS(base, x, y, adj, val, verbose=False, _features_override=feats)
    assert sorted(k for k in cal.state_dict() if k.startswith("net.")) == \
        ["net.0.bias", "net.0.weight", "net.2.bias", "net.2.weight"]
    assert "wavelet_feats" not in cal.state_dict()
    for k in ("net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias"):
        np.testing.assert_allclose(cal.state_dict()[k].numpy(), g[k], rtol=0, atol=1e-6)
    cal.eval()
    with torch.no_grad():
        lp = cal(x, adj).numpy()
    acc, conf, ece = orc.evaluate_probs(lp, y.numpy(), test.numpy())
    assert round(acc, 4) == round(float(g["acc"]), 4)
    assert round(conf, 4) == round(float(g["conf"]), 4)
    assert round(ece, 4) == round(float(g["ece"]), 4)


def test_laplacian_operator_algebra_matches_the_reference_expressions():
    """`(2.0 / lambda_max) * L - identity(N)` (calibration/WATS.py:55) and `2 * L` (:36) on the
    implicit operator: only the two scalars change; anything but a multiple of I is refused."""
    import scipy.sparse as sp
    from efficient_gnn_b200.wats import LaplacianOperator

    class FakeGraph:
        n = 5

    L = LaplacianOperator(FakeGraph())
    assert (L.scale, L.shift, L.shape) == (1.0, 0.0, (5, 5))
    Lr = (2.0 / 2.0) * L - sp.identity(5)
    assert (Lr.scale, Lr.shift) == (1.0, -1.0)
    L3 = (2.0 / 3.0) * L - sp.identity(5)
    assert L3.scale == pytest.approx(2.0 / 3.0) and L3.shift == -1.0
    twice = 2 * Lr
    assert (twice.scale, twice.shift) == (2.0, -2.0)
    same = L.rescaled(3.0)
    assert same.scale == pytest.approx(L3.scale) and same.shift == L3.shift
    plus = Lr + 0.5 * sp.identity(5)
    assert plus.shift == -0.5
    with pytest.raises(ValueError):
        L - sp.csr_matrix(np.ones((5, 5)))
    with pytest.raises(ValueError):
        L - sp.identity(4)
    with pytest.raises(TypeError):
        L - 1.0


def test_heat_coefficients_shapes_and_values():
    from efficient_gnn_b200.wats import heat_coefficients
    c = heat_coefficients(3, [0.4, 0.8, 1.6])
    assert c.shape == (3, 4) and c.dtype == np.float64
    np.testing.assert_allclose(c[:, 0], 1.0)
    np.testing.assert_allclose(c[1], np.exp(-0.8 * np.arange(4)))
    assert heat_coefficients(0, 0.8).shape == (1, 1)
