"""SURVEY 8c's downstream recipe, literally: the REFERENCE's own ``WATS`` class
(calibration/WATS.py:76-170, imported unmodified from oracle/_ref) is
constructed twice on the same device with the same seeds - once untouched
(scipy path on the host), once with ``calibration.WATS.graph_wavelet_features``
monkey-patched to the CUDA builder, exactly the one-line change INTEGRATION.md
asks a maintainer to make - and evaluated as
benchmark_calibration_methods.py:100-127 does (accuracy, mean max-probability,
the reference's ``calculate_average_ece``): equal to 4 decimals."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
from oracle import ref_shim
from models_for_tests import FixedLogits

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(ref_shim.reference_root() is None,
                                 reason="oracle/_ref not staged (run __graft_entry__.build() next to /root/reference)")]


def evaluate(ece_mod, model, x, y, adj, test_mask):
    """benchmark_calibration_methods.py:100-127 (evaluate_calibration) without the prints."""
    model.eval()
    with torch.no_grad():
        probs = model(x, adj).exp()
        tp, tl = probs[test_mask], y[test_mask]
        acc = (torch.argmax(tp, dim=1) == tl).float().mean().item()
        conf = torch.max(tp, dim=1)[0].mean().item()
        ece = ece_mod.calculate_average_ece(tp.cpu().numpy(), tl.cpu().numpy(), tp.shape[1], logits=False)
    return acc, conf, float(ece)


def build(wats_mod, model_mod, shape, self_loops, use_gcn):
    sh = synth.SHAPES[shape]
    rp, ci, n = synth.synth_csr(shape, self_loops=self_loops)
    adj_csr = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n))
    adj = torch.tensor(adj_csr.toarray(), dtype=torch.float32)
    y, logits, val, test = synth.synth_labels(sh.n, sh.n_classes, seed=42)
    x = torch.randn(sh.n, 16, generator=torch.Generator().manual_seed(7))
    torch.manual_seed(42)                       # benchmark_calibration_methods.py:166-167
    torch.cuda.manual_seed_all(42)
    np.random.seed(42)
    base = model_mod.CompatibleGCN(16, nclass=sh.n_classes) if use_gcn else FixedLogits(logits)
    cal = wats_mod.WATS(base, x, y, adj, val)   # the reference constructor: features + calib_train
    return cal, x.cuda(), y.cuda(), adj.cuda(), test.cuda()


@pytest.mark.parametrize("shape,self_loops,use_gcn", [("cora", False, True), ("cora", True, False),
                                                      ("pubmed", True, True)])
def test_reference_wats_class_with_the_cuda_feature_builder(monkeypatch, shape, self_loops, use_gcn):
    wats_mod, model_mod, ece_mod = ref_shim.load_reference()
    assert wats_mod.WATS.__module__ == "calibration.WATS" and wats_mod is not egnn.wats
    stock = build(wats_mod, model_mod, shape, self_loops, use_gcn)
    m_stock = evaluate(ece_mod, *stock)

    calls = []

    def cuda_builder(adj_matrix, k=3, s=0.8):
        calls.append((adj_matrix.shape, k, s))
        return egnn.graph_wavelet_features(adj_matrix, k, s)

    monkeypatch.setattr(wats_mod, "graph_wavelet_features", cuda_builder)
    swapped = build(wats_mod, model_mod, shape, self_loops, use_gcn)
    m_swapped = evaluate(ece_mod, *swapped)
    assert calls == [((synth.SHAPES[shape].n,) * 2, 3, 0.8)]          # WATS.py:99 reached the CUDA builder
    f_stock, f_swapped = stock[0].wavelet_feats, swapped[0].wavelet_feats
    assert f_swapped.is_cuda and f_swapped.dtype == torch.float32 and f_swapped.shape == f_stock.shape
    assert torch.allclose(f_swapped, f_stock, rtol=0, atol=1e-6)
    assert list(stock[0].state_dict()) == list(swapped[0].state_dict())
    for a, b, what in zip(m_swapped, m_stock, ("accuracy", "confidence", "ece")):
        assert round(a, 4) == round(b, 4), f"{what}: cuda {a} vs reference {b}"
    print(f"\n[{shape}] acc/conf/ece reference {m_stock} cuda-fed {m_swapped}")
