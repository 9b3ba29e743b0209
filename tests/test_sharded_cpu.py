"""Host logic of the row-sharded path on CPU: world_size 2 (and 3) over gloo,
with a NumPy engine standing in for the CUDA kernels.  What is under test is
the partition, the local/remote column split, the exchange order and the slab
rotation - the arithmetic here is the oracle's, so any mismatch is plumbing."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NumpyEngine:
    """CPU stand-in with the CudaEngine interface (tests only)."""

    def __init__(self):
        self.calls = []

    def prep(self, rowptr, colidx, n_global, row_begin, n_rows, allreduce):
        rp, ci = rowptr.numpy(), colidx.numpy()
        colsum = torch.zeros(n_global, dtype=torch.float64)
        diag = torch.zeros(n_global, dtype=torch.float32)
        rows = np.repeat(np.arange(n_rows), np.diff(rp)) + row_begin
        np.add.at(colsum.numpy(), ci, 1.0)
        on = rows == ci
        diag.numpy()[rows[on]] = 1.0
        rowsum = np.diff(rp).astype(np.float32)
        allreduce(colsum)
        allreduce(diag)
        w = colsum.numpy().astype(np.float32) - diag.numpy()
        iso = w == 0
        dinv = np.where(iso, 1.0, 1.0 / np.sqrt(np.where(iso, 1.0, w).astype(np.float64))).astype(np.float32)
        self.w_full, self.rowsum_local = w.copy(), rowsum.copy()
        return (torch.from_numpy(dinv), torch.from_numpy(iso.astype(np.uint8)),
                torch.from_numpy(np.log1p(rowsum)), False)

    def patch_degrees(self, dinv, iso, x0_local, n_global, row_begin, n_rows, deltas):
        """Host restatement of egnn_patch_degrees on a row shard."""
        w, rowsum = self.w_full.copy(), self.rowsum_local.copy()
        for r, c, v in zip(*deltas):
            if r != c:
                w[c] += v
            if row_begin <= r < row_begin + n_rows:
                rowsum[r - row_begin] += v
        is_iso = w == 0
        d2 = np.where(is_iso, 1.0, 1.0 / np.sqrt(np.where(is_iso, 1.0, w).astype(np.float64))).astype(np.float32)
        return (torch.from_numpy(d2), torch.from_numpy(is_iso.astype(np.uint8)),
                torch.from_numpy(np.log1p(rowsum).astype(np.float32)))

    def sell_plan(self, *a, **k):
        return self.fake_plan

    fake_plan = None

    def fork(self):
        self.calls.append("fork")

    def join(self):
        self.calls.append("join")

    def side_stream(self):
        import contextlib
        return contextlib.nullcontext()

    def order(self, phase, local, remote, dinv, iso, t_prev_full, t_prev_local, t_prev2_local, t_out_local,
              out_local, acc_ws, n_global, nnz_hint, row_begin, row_end, f, order, k_max, n_scales, coeffs,
              op_scale, op_shift, normalize, deltas=None):
        self.calls.append(f"order{order}.phase{phase}")
        rows = row_end - row_begin
        rp, ci = (local if phase == 0 else remote)
        rp, ci = rp.numpy(), ci.numpy()
        d = dinv.numpy().astype(np.float64)
        src = t_prev_local.numpy().astype(np.float64) if phase == 0 else t_prev_full.numpy().astype(np.float64)
        off = row_begin if phase == 0 else 0
        acc = np.zeros((rows, f))
        r = np.repeat(np.arange(rows), np.diff(rp))
        keep = ci != (r + row_begin)
        np.add.at(acc, r[keep], d[ci[keep], None] * src[ci[keep] - off])
        if deltas is not None:
            assert phase != 0                  # flips ride with the launch that sees the exchanged operand
            for dr, dc, dv in zip(*deltas):
                if row_begin <= dr < row_end and dr != dc:
                    acc[dr - row_begin] += dv * d[dc] * src[dc]
        if phase == 0:
            acc_ws.numpy()[:rows] = acc
            return
        acc += acc_ws.numpy()[:rows]
        di = d[row_begin:row_end, None]
        theta = (op_scale * (1 - iso.numpy()[row_begin:row_end]) + op_shift)[:, None]
        xprev = t_prev_local.numpy()[:rows].astype(np.float64)
        lap = theta * xprev - op_scale * di * acc
        tk = lap if order == 1 else 2 * lap - t_prev2_local.numpy()[:rows]
        if t_out_local is not None:
            t_out_local.numpy()[:rows] = tk
        o = out_local.numpy()
        for s in range(n_scales):
            ck, cp = coeffs[s, order], coeffs[s, order - 1]
            val = cp * xprev + ck * tk if order == 1 else o[:rows, s, :] + ck * tk
            if normalize and order == k_max:
                val = val / (np.abs(val).sum(axis=1, keepdims=True) + 1e-8)
            o[:rows, s, :] = val


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, f, k, scales, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from efficient_gnn_b200 import sharded, synth
        from oracle import wats_oracle as orc
        rp, ci, n = synth.synth_csr(synth.GraphShape("t", 1003, 9000, 3, 21, 1), self_loops=True)
        part = sharded.RowPartition(n, world)
        rpl, cil = part.slice_csr(rp, ci, rank)
        eng = NumpyEngine()
        sw = sharded.ShardedWavelet(rpl, cil, n, engine=eng, device="cpu")
        x0 = None
        x0_full = None
        if f > 1:
            x0_full = np.random.default_rng(9).standard_normal((n, f)).astype(np.float32)
            x0 = torch.from_numpy(x0_full[sw.row_begin:sw.row_end])
        feats, orders, comb = sw.features(k=k, s=scales, X0_local=x0, return_parts=True)
        fused = sw.features(k=k, s=scales, X0_local=x0)
        full = sw.gather_features(fused)
        adj = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n))
        p = orc.wavelet_parts(adj, k=k, s=scales, x0=x0_full)
        b, e = sw.row_begin, sw.row_end
        for got, ref in zip(orders, p["T"]):
            np.testing.assert_allclose(got.numpy(), ref[b:e], atol=2e-5)
        want = np.concatenate(p["H"], axis=1)
        np.testing.assert_allclose(feats.numpy(), want[b:e], atol=2e-5)
        np.testing.assert_allclose(full.numpy(), want, atol=2e-5)
        # exchange runs between the two halves of every order
        per_order = [c for c in eng.calls if c.startswith("order") or c in ("fork", "join")]
        assert per_order[:4] == ["fork", "order1.phase0", "join", "order1.phase1"]
        # halves partition the shard's entries
        assert sw.local_half[1].numel() + sw.remote_half[1].numel() == cil.numel()
        assert bool(((sw.local_half[1] >= b) & (sw.local_half[1] < e)).all())
        assert not bool(((sw.remote_half[1] >= b) & (sw.remote_half[1] < e)).any())
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,f,k,scales", [(2, 1, 3, 0.8), (2, 3, 4, [0.8, 1.6]), (3, 2, 2, 0.8)])
def test_sharded_host_logic_gloo(world, f, k, scales):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, f, k, scales, ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world


def test_row_partition_and_split():
    sys.path.insert(0, ROOT)
    from efficient_gnn_b200 import sharded
    part = sharded.RowPartition(10, 4)
    assert [part.begin(r) for r in range(4)] == [0, 3, 6, 9]
    assert [part.rows(r) for r in range(4)] == [3, 3, 3, 1]
    part = sharded.RowPartition(3, 8)                      # more ranks than rows
    assert sum(part.rows(r) for r in range(8)) == 3
    rowptr = torch.tensor([0, 2, 5, 5], dtype=torch.int32)
    colidx = torch.tensor([0, 7, 1, 2, 9], dtype=torch.int32)
    (lp, lc), (rp, rc) = sharded.split_columns(rowptr, colidx, 0, 3)
    assert lp.tolist() == [0, 1, 3, 3] and lc.tolist() == [0, 1, 2]
    assert rp.tolist() == [0, 1, 2, 2] and rc.tolist() == [7, 9]


def _delta_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from efficient_gnn_b200 import sharded, synth
        from oracle import wats_oracle as orc
        rp, ci, n = synth.synth_csr(synth.GraphShape("t", 901, 8000, 3, 33, 1), self_loops=True)
        adj = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n))
        dense = adj.toarray()
        # calib_fga.py:897-904: symmetric flips incident to one target node, across shard boundaries
        target, others = 5, [7, 300, 450, 700, 899]
        rows, cols, vals = [], [], []
        pert = dense.copy()
        for j in others:
            v = float(-2 * dense[target, j] + 1)
            pert[target, j] += v
            pert[j, target] += v
            rows += [target, j]; cols += [j, target]; vals += [v, v]
        part = sharded.RowPartition(n, world)
        rpl, cil = part.slice_csr(rp, ci, rank)
        sw = sharded.ShardedWavelet(rpl, cil, n, engine=NumpyEngine(), device="cpu")
        feats, orders, comb = sw.features(k=3, s=[0.8, 1.6], return_parts=True, deltas=(rows, cols, vals))
        p = orc.wavelet_parts(sp.csr_matrix(pert.astype(np.float32)), k=3, s=[0.8, 1.6])
        b, e = sw.row_begin, sw.row_end
        for got, ref in zip(orders, p["T"]):
            np.testing.assert_allclose(got.numpy(), ref[b:e], atol=2e-5)
        np.testing.assert_allclose(feats.numpy(), np.concatenate(p["H"], axis=1)[b:e], atol=2e-5)
        # and the unperturbed graph is untouched by the call
        base = sw.features(k=3, s=0.8)
        np.testing.assert_allclose(base.numpy(), orc.wavelet_features(adj, k=3, s=0.8)[b:e], atol=2e-5)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_edge_flips_gloo():
    """UGCA recompute on a row-sharded graph (BASELINE config 5): global flip list, every rank
    patches its replicated degree vectors and applies the flips of the rows it owns."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_delta_worker, args=(2, port, ret), nprocs=2, join=True)
    assert [ret.get(r) for r in range(2)] == ["ok"] * 2


def _mixed_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from efficient_gnn_b200 import sharded, synth
        rp, ci, n = synth.synth_csr(synth.GraphShape("t", 400, 3000, 3, 11, 1), self_loops=True)
        part = sharded.RowPartition(n, world)
        rpl, cil = part.slice_csr(rp, ci, rank)
        eng = NumpyEngine()
        eng.fake_plan = object() if rank == 0 else None      # only rank 0's shard qualifies for the SELL plan
        sw = sharded.ShardedWavelet(rpl, cil, n, engine=eng, device="cpu")
        ret[rank] = "no plan" if sw.plan is None else "plan"
        feats = sw.features(k=2, s=0.8)                       # and the generic path still runs on both
        assert feats.shape == (sw.rows, 1)
    finally:
        dist.destroy_process_group()


def test_ranks_agree_on_the_narrow_path_gloo():
    """A rank whose shard does not qualify for the SELL plan takes it away from every rank:
    the plan path exchanges dinv*T, the generic path plain T - a mixed world would silently
    read each other's slabs as the wrong quantity."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_mixed_worker, args=(2, port, ret), nprocs=2, join=True)
    assert [ret.get(r) for r in range(2)] == ["no plan", "no plan"]


def _skewed_csr(n=603, seed=3):
    """Degree-sorted ids (longest rows first): equal-rows shards are badly unbalanced."""
    rng = np.random.default_rng(seed)
    deg = np.sort(np.minimum((rng.random(n) ** -0.7).astype(np.int64) + 1, 120))[::-1]
    a = sp.lil_matrix((n, n), dtype=np.float32)
    for i, d in enumerate(deg):
        cols = rng.choice(n, size=int(d), replace=False)
        a[i, cols] = 1.0
        a[cols, i] = 1.0
    a = sp.csr_matrix(a)
    a.sort_indices()
    return a


@pytest.mark.parametrize("world", [2, 3, 8])
def test_balanced_order_equalises_entries_and_keeps_the_features(world):
    """SURVEY 8e (nnz-balanced shards): the relabelling leaves every rank the same number of
    entries within one long row, fills every id range of the partition exactly, and the features
    of the relabelled graph are the original ones with permuted rows."""
    from efficient_gnn_b200 import sharded
    from oracle import wats_oracle as orc
    adj = _skewed_csr()
    n = adj.shape[0]
    rp = torch.from_numpy(adj.indptr.astype(np.int32))
    ci = torch.from_numpy(adj.indices.astype(np.int32))
    before = sharded.BalancedOrder.shard_entries(rp, world)
    order = sharded.BalancedOrder.from_rowptr(rp, world)
    assert sorted(order.perm.tolist()) == list(range(n))
    rp2, ci2, _ = order.relabel_csr(rp, ci)
    after = sharded.BalancedOrder.shard_entries(rp2, world)
    assert sum(after) == sum(before) == adj.nnz
    assert max(before) > 1.5 * min(before)
    assert max(after) - min(after) <= int(np.diff(adj.indptr).max())
    # rows stay column-sorted and the graph is the same up to the renumbering
    adj2 = sp.csr_matrix((np.ones(adj.nnz, np.float32), ci2.numpy(), rp2.numpy()), shape=(n, n))
    assert adj2.has_sorted_indices or (adj2.sort_indices() is None and np.array_equal(adj2.indices, ci2.numpy()))
    p = order.perm.numpy()
    assert (adj2 != adj[p][:, p]).nnz == 0
    want = orc.wavelet_parts(adj, k=3, s=0.8)["S"][0]               # un-normalised (the F = 1 feature itself is a sign)
    got = orc.wavelet_parts(adj2, k=3, s=0.8)["S"][0]
    np.testing.assert_allclose(order.to_original(torch.from_numpy(np.asarray(got))).numpy(), want, atol=1e-12)
    # ids going in (edge flips, signals) follow the same map
    rows, cols, vals = order.relabel_deltas(([5, 9], [9, 5], [1.0, 1.0]))
    assert p[rows[0]] == 5 and p[cols[0]] == 9 and vals == [1.0, 1.0]
    x = torch.arange(n, dtype=torch.float32)
    assert torch.equal(order.to_original(order.from_original(x)), x)


def test_balanced_order_fills_short_last_ranges():
    from efficient_gnn_b200 import sharded
    for n, world in [(10, 4), (7, 8), (1, 3), (16, 4), (9, 2)]:
        rp = torch.arange(n + 1, dtype=torch.int32) * 3
        order = sharded.BalancedOrder.from_rowptr(rp, world)
        assert sorted(order.perm.tolist()) == list(range(n))
        weighted = torch.arange(n, 0, -1, dtype=torch.int64)          # with values as well
        rp2, ci2, v2 = order.relabel_csr(torch.cat([torch.zeros(1, dtype=torch.int64), weighted.cumsum(0)]).int(),
                                         torch.cat([torch.arange(d) % n for d in weighted.tolist()]).int(),
                                         torch.arange(int(weighted.sum()), dtype=torch.float32))
        assert int(rp2[-1]) == int(weighted.sum()) and v2.numel() == ci2.numel()


def _balanced_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from efficient_gnn_b200 import sharded
        from oracle import wats_oracle as orc
        adj = _skewed_csr(n=402, seed=11)
        n = adj.shape[0]
        rp = torch.from_numpy(adj.indptr.astype(np.int32))
        ci = torch.from_numpy(adj.indices.astype(np.int32))
        order = sharded.BalancedOrder.from_rowptr(rp, world)          # same on every rank (deterministic)
        rp2, ci2, _ = order.relabel_csr(rp, ci)
        part = sharded.RowPartition(n, world)
        rpl, cil = part.slice_csr(rp2, ci2, rank)
        sw = sharded.ShardedWavelet(rpl, cil, n, engine=NumpyEngine(), device="cpu")
        # un-normalised combination (the F = 1 feature itself is a sign): original ids in, original order out
        local = sw.features(k=3, s=0.8, normalize=False)
        got = order.to_original(sw.gather_features(local))
        want = orc.wavelet_parts(adj, k=3, s=0.8)["S"][0]
        np.testing.assert_allclose(got.numpy(), want, atol=2e-5 * np.abs(want).max())
        # an edge flip named by ORIGINAL node ids goes through the same map
        u, v = 3, 250
        val = float(1 - 2 * adj[u, v])
        pert = adj.toarray()
        pert[u, v] += val
        pert[v, u] += val
        flips = order.relabel_deltas(([u, v], [v, u], [val, val]))
        got_d = order.to_original(sw.gather_features(sw.features(k=3, s=0.8, normalize=False, deltas=flips)))
        want_d = orc.wavelet_parts(sp.csr_matrix(pert.astype(np.float32)), k=3, s=0.8)["S"][0]
        np.testing.assert_allclose(got_d.numpy(), want_d, atol=2e-5 * np.abs(want_d).max())
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_balanced_renumbering_end_to_end_gloo():
    """Skewed numbering -> BalancedOrder -> row shards over gloo -> features mapped back: equal to
    the oracle on the ORIGINAL graph, with and without an edge flip given in original ids."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_balanced_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert [ret.get(r) for r in range(world)] == ["ok"] * world
