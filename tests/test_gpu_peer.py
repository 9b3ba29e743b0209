"""The fused exchange over peer memory (csrc/peer.cuh) needs one process per
GPU on >= 2 real GPUs: spin-waiting kernels of different ranks must never share
a device.  On a single-GPU box only the one-rank window (no peers, plain
stores) is exercised; the 2-rank cases run under ``gpurun --gpus 2``."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from efficient_gnn_b200 import sharded, synth
from oracle import wats_oracle as orc

pytestmark = pytest.mark.gpu

SHAPE = synth.GraphShape("t", 6001, 260_000, 3, 33, 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flips(n):
    """Symmetric flips around one target node, spread over both shards (calib_fga.py:897-904)."""
    target, others = 11, [3, 17, n // 2 + 5, n - 2, n // 3]
    return target, others


def _worker(rank, world, port, k, scales, steps, out_dir, f=1, flips=False):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        rp, ci, n = synth.synth_csr(SHAPE, self_loops=True)
        part = sharded.RowPartition(n, world)
        rpl, cil = part.slice_csr(rp, ci, rank)
        sw = sharded.ShardedWavelet(rpl.to(dev), cil.to(dev), n, device=dev)
        assert sw.plan is not None and sw.peer is not None and sw.fused_wide, "fused exchange was not selected"
        x0 = None
        if f != 1:                                   # f == -1: a custom one-column signal (order-1 operand is exchanged too)
            x0_full = np.random.default_rng(abs(f)).standard_normal((n, abs(f))).astype(np.float32)
            x0 = torch.from_numpy(x0_full[sw.row_begin:sw.row_end]).to(dev)
        deltas = None
        if flips:                                    # UGCA recompute on the sharded graph: global ids, same list on every rank
            dense_row = torch.zeros(n)
            target, others = _flips(n)
            dense_row[ci[rp[target]:rp[target + 1]].long()] = 1.0
            rows, cols, vals = [], [], []
            for j in others:
                v = float(-2 * dense_row[j].item() + 1)
                rows += [target, j]; cols += [j, target]; vals += [v, v]
            deltas = (rows, cols, vals)
        outs = []
        for step in range(steps):                    # consecutive steps reuse the two operand buffers
            outs.append(sw.features(k=k, s=scales, X0_local=x0, deltas=deltas).cpu().numpy())
            if flips and step == 0:                  # perturbed and unperturbed steps interleave in the UGCA loop
                sw.features(k=k, s=scales, X0_local=x0)
        feats, orders, comb = sw.features(k=k, s=scales, X0_local=x0, return_parts=True, deltas=deltas)
        sw.check_exchange()                          # raises if a flag wait timed out
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), b=sw.row_begin, e=sw.row_end,
                 fused=np.stack(outs), comb=comb.cpu().numpy(), **{f"t{i}": o.cpu().numpy() for i, o in enumerate(orders)})
        sw.close()
        assert sw.peer is None and not sw._wide_peers
    finally:
        dist.destroy_process_group()


# steps > 3 with even K: the buffer the first producer of a step fills was read by the peers'
# last order of the previous step (the wait-before-first-store of the step kernel)
@pytest.mark.parametrize("k,scales,f,steps,flips", [
    (3, 0.8, 1, 3, False), (4, [0.8, 1.6], 1, 12, False), (6, 0.8, 1, 8, False), (2, 0.8, 1, 8, False),
    (1, 0.8, 1, 3, False), (3, [0.8, 1.6], -1, 3, False), (4, 0.8, -1, 8, False),
    (3, [0.8, 1.6], 16, 3, False), (4, 0.8, 130, 3, False),
    (3, 0.8, 1, 3, True), (4, [0.8, 1.6], 1, 4, True), (3, 0.8, 16, 3, True),
    (18, 0.8, 1, 3, False), (19, [0.8, 1.6], 1, 3, True)])                  # K > 16: two launches per step
def test_two_rank_fused_exchange_matches_oracle(tmp_path, k, scales, f, steps, flips):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one process per GPU)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), k, scales, steps, str(tmp_path), f, flips), nprocs=world, join=True)
    rp, ci, n = synth.synth_csr(SHAPE, self_loops=True)
    adj = sp.csr_matrix((np.ones(ci.numel(), np.float32), ci.numpy(), rp.numpy()), shape=(n, n))
    if flips:
        dense = adj.toarray()
        target, others = _flips(n)
        for j in others:
            v = -2 * dense[target, j] + 1
            dense[target, j] += v
            dense[j, target] += v
        adj = sp.csr_matrix(dense.astype(np.float32))
    x0_full = None if f == 1 else np.random.default_rng(abs(f)).standard_normal((n, abs(f))).astype(np.float32)
    p = orc.wavelet_parts(adj, k=k, s=scales, x0=x0_full)
    want_h = np.concatenate(p["H"], axis=1).astype(np.float32)
    want_s = np.stack(p["S"], axis=1)                                   # [N, S, F]
    sure = np.concatenate([np.abs(sj) > 1e-4 * np.abs(sj).max() for sj in p["S"]], axis=1)
    tol = 1e-5 if k <= 6 else 1e-4                                      # fp32 recurrence over many orders
    for rank in range(world):
        z = np.load(tmp_path / f"rank{rank}.npz")
        b, e = int(z["b"]), int(z["e"])
        for i in range(k + 1):
            ref = p["T"][i]
            assert np.abs(z[f"t{i}"] - ref[b:e]).max() / np.abs(ref).max() <= tol, f"order {i}"
        assert np.abs(z["comb"] - want_s[b:e]).max() / np.abs(want_s).max() <= tol
        for step in range(steps):
            got = z["fused"][step]
            np.testing.assert_allclose(got[sure[b:e]], want_h[b:e][sure[b:e]], atol=2e-5)
        assert np.array_equal(z["fused"][0], z["fused"][-1])           # deterministic across steps


def _alternating_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        rp, ci, n = synth.synth_csr(SHAPE, self_loops=True)
        assert n % 4 != 0                             # the last column block ends in a partial 16-byte group
        part = sharded.RowPartition(n, world)
        rpl, cil = part.slice_csr(rp, ci, rank)
        sw = sharded.ShardedWavelet(rpl.to(dev), cil.to(dev), n, device=dev)
        x_full = np.random.default_rng(5).standard_normal((n, 1)).astype(np.float32)
        x = torch.from_numpy(x_full[sw.row_begin:sw.row_end]).to(dev)
        bad = 0
        for k in (3, 4):
            base = sw.features(k=k, s=0.8, X0_local=x, normalize=False).clone()
            for step in range(10):                    # every step stages different operand values
                scale = float(2 ** (step % 3 + 1))
                got = sw.features(k=k, s=0.8, X0_local=scale * x, normalize=False)
                bad += int(not torch.equal(got, scale * base))      # scaling by a power of two is exact in float32
        sw.check_exchange()
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write(str(bad))
        sw.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_consecutive_steps_never_see_stale_operands(tmp_path):
    """Steps whose operands differ (a power-of-two multiple of the same signal: results must be
    the exact multiple): any element staged before its owner's flag arrived - e.g. the last,
    partial 16-byte group of a column block, which is not part of the bulk copy - shows up."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one process per GPU)")
    import torch.multiprocessing as mp
    mp.spawn(_alternating_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"rank{r}.txt").read() for r in range(2)] == ["0", "0"]


def _lopsided_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    import efficient_gnn_b200 as egnn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        # directed graph over 4 column blocks: the rows of rank 0 are long and point everywhere, the
        # rows of rank 1 are short and point only at rank 1's own columns
        cb = 49152
        n, half, deg = 4 * cb, 2 * cb, 600
        top = (np.arange(half, dtype=np.int64)[:, None] * 7 + np.arange(deg, dtype=np.int64)[None, :] * 327) % n
        top.sort(axis=1)
        low_deg = 40                                  # enough for the plan (8 entries per row and column block)
        low = half + (np.arange(half, dtype=np.int64)[:, None] * 5 + np.arange(low_deg, dtype=np.int64)[None, :] * 1201) % half
        low.sort(axis=1)
        ci = torch.from_numpy(np.concatenate([top.ravel(), low.ravel()]).astype(np.int32))
        rp = torch.from_numpy(np.concatenate([np.arange(half + 1, dtype=np.int64) * deg,
                                              half * deg + low_deg * np.arange(1, half + 1, dtype=np.int64)]).astype(np.int32))
        part = sharded.RowPartition(n, world)
        assert part.rows_per == half
        rpl, cil = part.slice_csr(rp, ci, rank)
        sw = sharded.ShardedWavelet(rpl.to(dev), cil.to(dev), n, device=dev)
        assert sw.plan is not None and sw.peer is not None
        x_full = np.random.default_rng(9).standard_normal((n, 1)).astype(np.float32)
        x = torch.from_numpy(x_full[sw.row_begin:sw.row_end]).to(dev)
        g = egnn.CsrGraph(rp.to(dev), ci.to(dev), None, n)
        want = egnn.graph_wavelet_features(g, k=8, s=0.8, X0=torch.from_numpy(x_full).to(dev), normalize=False)
        want = want[sw.row_begin:sw.row_end]
        worst = 0.0
        for step in range(6):
            got = sw.features(k=8, s=0.8, X0_local=x, normalize=False)
            worst = max(worst, float((got - want).abs().max() / want.abs().max()))
        sw.check_exchange()
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write(repr(worst))
        sw.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_light_shard_never_overwrites_a_window_still_being_read(tmp_path):
    """A rank whose slices touch none of a slower rank's columns is not held back by any operand
    wait; before it stores the next operand into that rank's window it must still know that the
    rank has finished staging the order that read the same buffer (K >= 4 reuses each buffer)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one process per GPU)")
    import torch.multiprocessing as mp
    mp.spawn(_lopsided_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    worst = [float(open(tmp_path / f"rank{r}.txt").read()) for r in range(2)]
    assert max(worst) <= 1e-5, worst


def test_single_rank_window_is_a_plain_store(tmp_path):
    """world == 1: same kernels, no peers, no waits - must equal the NCCL-free path."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()), RANK="0", WORLD_SIZE="1")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        rp, ci, n = synth.synth_csr(SHAPE, self_loops=True)
        plain = sharded.ShardedWavelet(rp.to(dev), ci.to(dev), n, device=dev, peer_exchange=False)
        peer = sharded.PeerExchange(plain.part.rows_per, 1, device=dev)
        fused = sharded.ShardedWavelet(rp.to(dev), ci.to(dev), n, device=dev, peer_exchange=peer)
        a = plain.features(k=3, s=[0.8, 1.6])
        b = fused.features(k=3, s=[0.8, 1.6])
        c = fused.features(k=3, s=[0.8, 1.6])
        assert peer.error() == 0
        assert torch.equal(a, b) and torch.equal(b, c)
        peer.close()
    finally:
        dist.destroy_process_group()
