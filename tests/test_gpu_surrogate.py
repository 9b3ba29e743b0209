"""the sparse structure-gradient surrogate against torch autograd on the
DENSE model, the way the attack computes it (calib_attack/calib_fga.py:864-890):
``output = surrogate(x, adj_leaf)[[target]]``, ``grad = autograd.grad(loss, adj_leaf)``,
of which only ``grad[target]`` and ``grad[:, target]`` are read (:881).  The base model
is the reference's two-layer row-normalised GCN (src/gnn/model.py:43-52; restated in
tests/models_for_tests.DenseGCN, and taken from oracle/_ref when staged)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.nn.functional as F

import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
from oracle import ref_shim
from models_for_tests import DenseGCN

pytestmark = pytest.mark.gpu


def make_problem(n, nnz, self_loops, seed=3, f=50, c=7, weighted=False):
    rp, ci, n = synth.synth_csr(synth.GraphShape("t", n, nnz, c, seed, 1), self_loops=self_loops)
    data = np.ones(ci.numel(), np.float32)
    if weighted:
        data = np.random.default_rng(seed).uniform(0.5, 2.0, ci.numel()).astype(np.float32)
    adj = sp.csr_matrix((data, ci.numpy(), rp.numpy()), shape=(n, n))
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, f, generator=g)
    torch.manual_seed(seed)
    if ref_shim.reference_root() is not None:          # the reference's own CompatibleGCN when available
        gcn = ref_shim.load_reference()[1].CompatibleGCN(f, nclass=c)
    else:
        gcn = DenseGCN(f, c)
    gcn = gcn.cuda().eval()
    return adj, x.cuda(), gcn


def kl_to_uniform(out):
    """Calibration-style loss on the target's output (stand-in for calib_attack_loss.py's)."""
    logp = F.log_softmax(out, dim=1)
    return F.kl_div(logp, torch.full_like(logp, 1.0 / logp.shape[1]), reduction="batchmean")


def dense_reference(gcn, x, dense, target, loss_fn, temp=None):
    leaf = dense.clone().detach().requires_grad_(True)
    out = gcn(x, leaf)[[target]]
    if temp is not None:
        out = out / temp
    loss = loss_fn(out)
    grad = torch.autograd.grad(loss, leaf)[0]
    return out.detach(), grad[target].clone(), grad[:, target].clone()


def check(sur, gcn, x, dense, target, deltas=None, temp=None, loss_fn=kl_to_uniform, tol=2e-5):
    out_ref, row_ref, col_ref = dense_reference(gcn, x, dense, target, loss_fn, temp)
    out = sur.target_logits(target, deltas)
    if temp is not None:
        out = out / temp
    sg = sur.structure_gradient(loss_fn(out))
    assert torch.allclose(out.detach(), out_ref, rtol=1e-5, atol=1e-5)
    scale = max(row_ref.abs().max().item(), col_ref.abs().max().item(), 1e-12)
    assert (sg.row - row_ref).abs().max().item() <= tol * scale, "row"
    assert (sg.col - col_ref).abs().max().item() <= tol * scale, "column"
    # what the attack ranks flips by (calib_fga.py:880-881), and its argmax
    score_ref = (row_ref + col_ref) * (-2 * dense[target] + 1)
    score = sg.flip_scores(dense[target])
    assert (score - score_ref).abs().max().item() <= 2 * tol * scale


@pytest.mark.parametrize("self_loops,weighted", [(True, False), (False, False), (True, True)])
def test_row_and_column_gradient_match_dense_autograd(self_loops, weighted):
    adj, x, gcn = make_problem(3000, 40_000, self_loops, weighted=weighted)
    dense = torch.tensor(adj.toarray(), dtype=torch.float32, device="cuda")
    sur = egnn.SparseGCNSurrogate(gcn, x, adj)
    deg = np.asarray(adj.sum(axis=1)).ravel()
    targets = [0, 17, int(deg.argmax()), int(deg.argmin()), 2999]
    for t in targets:
        check(sur, gcn, x, dense, t)
        check(sur, gcn, x, dense, t, loss_fn=lambda out: -F.log_softmax(out, dim=1)[0, 2], temp=1.7)   # class loss on a scaled output
    # the full forward through the same propagation kernel
    with torch.no_grad():
        assert torch.allclose(sur.all_logits(), gcn(x, dense), rtol=1e-5, atol=1e-5)


def test_isolated_and_directed_targets():
    """Rows whose degree is clamped to 1 take no gradient through the degree
    (the in-place ``deg[deg == 0] = 1`` of src/gnn/model.py:49)."""
    n = 600
    rng = np.random.default_rng(1)
    dense_np = (rng.random((n, n)) < 0.01).astype(np.float32)       # directed, no symmetry
    dense_np[5, :] = 0                                              # node 5: no out-edges (clamped), but in-edges
    dense_np[:, 9] = 0
    dense_np[9, :] = 0                                              # node 9: isolated
    dense_np[7, 5] = 1                                              # 7 -> 5: a neighbour with a clamped row
    adj = sp.csr_matrix(dense_np)
    x = torch.randn(n, 20, generator=torch.Generator().manual_seed(2)).cuda()
    torch.manual_seed(2)
    gcn = DenseGCN(20, 4, nhid=32).cuda().eval()
    dense = torch.from_numpy(dense_np).cuda()
    sur = egnn.SparseGCNSurrogate(gcn, x, adj)
    for t in (5, 9, 7, 100):
        check(sur, gcn, x, dense, t)


def test_edge_flips_of_a_running_attack():
    """The perturbed adjacency of iteration i = base graph + i symmetric flips around the target
    (calib_fga.py:897-904), including the removal of an existing edge."""
    adj, x, gcn = make_problem(3000, 40_000, True)
    dense = torch.tensor(adj.toarray(), dtype=torch.float32, device="cuda")
    sur = egnn.SparseGCNSurrogate(gcn, x, adj)
    target = 17
    nbr = int(np.nonzero(adj[target].toarray().ravel())[0][1])
    others = [nbr, 5, 2500, 1234]
    rows, cols, vals = [], [], []
    pert = dense.clone()
    for j in others:
        v = float(-2 * pert[target, j].item() + 1)
        pert[target, j] += v
        pert[j, target] += v
        rows += [target, j]; cols += [j, target]; vals += [v, v]
        check(sur, gcn, x, pert, target, deltas=(list(rows), list(cols), list(vals)))
    check(sur, gcn, x, pert, others[1], deltas=(rows, cols, vals))         # a flipped partner as the target
    check(sur, gcn, x, dense, target)                                      # and the base graph is untouched


def test_reference_size_cap_20000_nodes():
    """The reference subsamples every graph to 20,000 nodes because of the dense tensors
    (exp/ablation/ugca_full_multi_dataset.py:575-579): parity at that size."""
    adj, x, gcn = make_problem(20_000, 400_000, True, seed=9, f=32, c=41)
    dense = torch.tensor(adj.toarray(), dtype=torch.float32, device="cuda")
    sur = egnn.SparseGCNSurrogate(gcn, x, adj)
    for t in (3, 19_999):
        check(sur, gcn, x, dense, t)


def test_three_gradients_from_one_forward():
    """calib_fga.py:877,889-890 takes three gradients (loss, p_max, p_2nd) of the same output."""
    adj, x, gcn = make_problem(1500, 20_000, True)
    dense = torch.tensor(adj.toarray(), dtype=torch.float32, device="cuda")
    sur = egnn.SparseGCNSurrogate(gcn, x, adj)
    t = 42
    leaf = dense.clone().requires_grad_(True)
    out_d = gcn(x, leaf)[[t]]
    out_s = sur.target_logits(t)
    for pick in (lambda o: kl_to_uniform(o), lambda o: torch.topk(F.softmax(o, 1), 2, dim=1)[0][0][0],
                 lambda o: torch.topk(F.softmax(o, 1), 2, dim=1)[0][0][1]):
        gd = torch.autograd.grad(pick(out_d), leaf, retain_graph=True)[0]
        sg = sur.structure_gradient(pick(out_s), retain_graph=True)
        scale = max(gd[t].abs().max().item(), 1e-12)
        assert (sg.row - gd[t]).abs().max().item() <= 2e-5 * scale
        assert (sg.col - gd[:, t]).abs().max().item() <= 2e-5 * scale
