"""Rank 0's share of a W-way row partition, run alone on one GPU with a no-op
exchange (timing only - the operand of the other ranks is garbage)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from efficient_gnn_b200 import sharded, synth
from efficient_gnn_b200.wats import WaveletSession

class NoComm:
    def __init__(self, rank, world): self.rank, self.world = rank, world
    def allreduce(self, t): pass
    def allgather(self, full, slab): full[:slab.shape[0]].copy_(slab)

world = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
rp, ci, n = synth.synth_csr("reddit", self_loops=True, device=dev)
part = sharded.RowPartition(n, world)
rpl, cil = part.slice_csr(rp, ci, 0)
del rp, ci
sw = sharded.ShardedWavelet(rpl, cil, n, device=dev, comm=NoComm(0, world))
print("plan", sw.plan is not None, "slices", sw.plan.n_slices, "entries", sw.plan.n_entries, "rowv", sw.plan.n_rowv, "blocks", sw.plan.n_blocks)
for graph in (False, True):
    ses = WaveletSession(sw, k=3, s=0.8, cuda_graph=graph)
    for _ in range(5): ses()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): ses()
    b.record(); torch.cuda.synchronize()
    print(f"world {world} graph={graph}: {a.elapsed_time(b)/steps*1e3:.1f} us/step", flush=True)
