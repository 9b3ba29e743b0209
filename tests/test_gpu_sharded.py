"""The row-sharded CUDA path on ONE GPU: the ranks of a 1/2/3/4-way partition
run as host threads sharing the device, with an in-process exchange standing
in for NCCL (the NCCL wiring itself is covered by tests/test_sharded_cpu.py
over gloo and by the multi-GPU bench).  Checks row offsets (row0 != 0), the
local/remote split kernels, the sharded SELL plan and the degree all-reduce."""
import threading

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import sharded, synth
from oracle import wats_oracle as orc
from helpers import rel_max_err

pytestmark = pytest.mark.gpu


class ThreadComm:
    """All ranks live in this process; collectives meet at a barrier.  The
    ranks share one device but launch from different threads and side streams,
    so each collective synchronises the device on both sides (inputs produced
    before anyone reads them, outputs complete before anyone overwrites)."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self.sh = rank, world, shared

    def allreduce(self, t):
        sh = self.sh
        sh["parts"][self.rank] = t
        torch.cuda.synchronize()
        sh["barrier"].wait()
        if self.rank == 0:
            total = torch.stack(sh["parts"]).sum(dim=0)
            for p in sh["parts"]:
                p.copy_(total)
            torch.cuda.synchronize()
        sh["barrier"].wait()

    def allgather(self, full, slab):
        sh = self.sh
        sh["slabs"][self.rank] = slab
        torch.cuda.synchronize()
        sh["barrier"].wait()
        rows = slab.shape[0]
        for r in range(self.world):
            full[r * rows:(r + 1) * rows].copy_(sh["slabs"][r])
        torch.cuda.synchronize()
        sh["barrier"].wait()


def run_world(world, rp, ci, n, f, k, scales, use_sell, x0_full, deltas=None):
    shared = {"barrier": threading.Barrier(world), "parts": [None] * world, "slabs": [None] * world}
    results, errors = [None] * world, []

    def work(rank):
        try:
            torch.cuda.set_device(0)
            part = sharded.RowPartition(n, world)
            rpl, cil = part.slice_csr(rp, ci, rank)
            sw = sharded.ShardedWavelet(rpl, cil, n, device="cuda", comm=ThreadComm(rank, world, shared),
                                        use_sell=use_sell)
            if use_sell:
                assert sw.plan is not None and sw.plan.row0 == part.begin(rank)
            x0 = None if x0_full is None else torch.from_numpy(x0_full[sw.row_begin:sw.row_end]).cuda()
            feats, orders, comb = sw.features(k=k, s=scales, X0_local=x0, return_parts=True, deltas=deltas)
            fused = sw.features(k=k, s=scales, X0_local=x0, deltas=deltas)
            results[rank] = (sw.row_begin, sw.row_end, [o.cpu().numpy() for o in orders], feats.cpu().numpy(),
                             fused.cpu().numpy(), sw.gather_features(fused).cpu().numpy())
        except Exception as exc:          # surface in the main thread, do not deadlock the barrier
            errors.append(exc)
            shared["barrier"].abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


@pytest.mark.parametrize("world,f,use_sell", [(1, 1, True), (2, 1, True), (4, 1, True), (2, 1, False),
                                              (3, 8, False), (2, 130, False)])
def test_sharded_matches_oracle(world, f, use_sell):
    shape = synth.GraphShape("t", 6001, 260_000, 3, 33, 1)
    rp, ci, n = synth.synth_csr(shape, self_loops=True)
    adj = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n))
    x0_full = None if f == 1 else np.random.default_rng(f).standard_normal((n, f)).astype(np.float32)
    k, scales = 3, [0.8, 1.6]
    p = orc.wavelet_parts(adj, k=k, s=scales, x0=x0_full)
    want = np.concatenate(p["H"], axis=1)
    res = run_world(world, rp.cuda(), ci.cuda(), n, f, k, scales, use_sell, x0_full)
    sure = np.concatenate([np.abs(sj) > 1e-4 * np.abs(sj).max() for sj in p["S"]], axis=1)
    for b, e, orders, feats, fused, gathered in res:
        for got, ref in zip(orders, p["T"]):
            assert np.abs(got - ref[b:e]).max() / np.abs(ref).max() <= 1e-5
        m = sure[b:e]
        np.testing.assert_allclose(feats[m], want[b:e][m], atol=2e-5)
        np.testing.assert_allclose(fused[m], want[b:e][m], atol=2e-5)
        np.testing.assert_allclose(gathered[sure], want[sure], atol=2e-5)


def test_single_rank_sharded_equals_single_gpu_path():
    rp, ci, n = synth.synth_csr("pubmed", self_loops=True)
    g = egnn.CsrGraph(rp.cuda(), ci.cuda(), None, n)
    a = egnn.graph_wavelet_features(g, k=3, s=0.8, return_parts=True, _use_sell=False)
    sw = sharded.ShardedWavelet(rp.cuda(), ci.cuda(), n, device="cuda", use_sell=False)
    feats, orders, comb = sw.features(k=3, s=0.8, return_parts=True)
    for x, y in zip(a.orders, orders):
        assert rel_max_err(y.cpu().numpy(), x.cpu().numpy()) <= 2e-6
    assert torch.equal(sw.dinv, g.dinv) and torch.equal(sw.iso, g.iso) and torch.equal(sw.x0, g.x0)


@pytest.mark.parametrize("world,f,use_sell", [(2, 1, True), (3, 1, True), (2, 1, False), (2, 8, False)])
def test_sharded_edge_flips_match_oracle(world, f, use_sell):
    """UGCA recompute on a row-sharded graph (BASELINE config 5): flips with global ids on top of
    the shards' CSR / SELL plans, against the oracle on the rebuilt adjacency."""
    shape = synth.GraphShape("t", 6001, 260_000, 3, 33, 1)
    rp, ci, n = synth.synth_csr(shape, self_loops=True)
    dense = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n)).toarray()
    target, others = 11, [3, 17, n // 2 + 5, n - 2, n // 3]
    rows, cols, vals = [], [], []
    for j in others:
        v = float(-2 * dense[target, j] + 1)
        dense[target, j] += v
        dense[j, target] += v
        rows += [target, j]; cols += [j, target]; vals += [v, v]
    x0_full = None if f == 1 else np.random.default_rng(f).standard_normal((n, f)).astype(np.float32)
    k, scales = 3, [0.8, 1.6]
    p = orc.wavelet_parts(sp.csr_matrix(dense), k=k, s=scales, x0=x0_full)
    res = run_world(world, rp.cuda(), ci.cuda(), n, f, k, scales, use_sell, x0_full, deltas=(rows, cols, vals))
    want = np.concatenate(p["H"], axis=1)
    sure = np.concatenate([np.abs(sj) > 1e-4 * np.abs(sj).max() for sj in p["S"]], axis=1)
    for b, e, orders, feats, fused, gathered in res:
        for got, ref in zip(orders, p["T"]):
            assert np.abs(got - ref[b:e]).max() / np.abs(ref).max() <= 1e-5
        m = sure[b:e]
        np.testing.assert_allclose(fused[m], want[b:e][m], atol=2e-5)
