"""Per-order cost of the fused exchange: steps with K = 0..3 orders on the Reddit shape (torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from efficient_gnn_b200 import sharded, synth
from efficient_gnn_b200.wats import WaveletSession
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rp, ci, n = synth.synth_csr("reddit", self_loops=True, device=dev)
part = sharded.RowPartition(n, world)
rpl, cil = part.slice_csr(rp, ci, rank); del rp, ci
sw = sharded.ShardedWavelet(rpl, cil, n, device=dev)
nnz_local = torch.tensor([float(cil.numel())], device=dev); allnnz = [torch.zeros_like(nnz_local) for _ in range(world)]
dist.all_gather(allnnz, nnz_local)
if rank == 0: print("nnz per rank", [int(t.item()) for t in allnnz], flush=True)
for k in (3,):
    ses = WaveletSession(sw, k=k, s=0.8, cuda_graph=True)
    for _ in range(10): ses()
    torch.cuda.synchronize(); dist.barrier()
    sw.peer.wait_stats()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): ses()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 200 * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ns, cnt = sw.peer.wait_stats()
    print(f"rank {rank} world {world} K={k}: {t.item():.1f} us/step; flag waits {cnt}, mean {ns / max(1, cnt) / 1e3:.2f} us", flush=True)
    del ses
print("err", sw.exchange_error()) if rank == 0 else None
dist.barrier(); os._exit(0)
