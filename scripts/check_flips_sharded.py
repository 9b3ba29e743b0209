"""Sharded edge flips at Reddit scale vs the single-GPU path (run under torchrun, N ranks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth, sharded
from bench import make_flips
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
rp, ci, n = synth.synth_csr(wl, self_loops=True, device=dev)
part = sharded.RowPartition(n, world)
rpl, cil = part.slice_csr(rp, ci, rank)
sw = sharded.ShardedWavelet(rpl, cil, n, device=dev)
sn = sharded.ShardedWavelet(rpl, cil, n, device=dev, peer_exchange=False)
g = egnn.CsrGraph(rp, ci, None, n)
b, e = sw.row_begin, sw.row_end
for seed in (100, 101, 102):
    d = make_flips(n, 5, seed)
    res = {}
    res["single parts"] = egnn.graph_wavelet_features(g, k=3, s=0.8, deltas=d, return_parts=True, _use_sell=True).combined[b:e, 0]
    res["single fused"] = egnn.graph_wavelet_features(g, k=3, s=0.8, deltas=d, normalize=False)[b:e]
    res["single generic"] = egnn.graph_wavelet_features(g, k=3, s=0.8, deltas=d, normalize=False, _use_sell=False)[b:e]
    res["peer parts"] = sw.features(k=3, s=0.8, deltas=d, return_parts=True)[2][:, 0]
    res["peer fused"] = sw.features(k=3, s=0.8, deltas=d, normalize=False)
    res["peer fused again"] = sw.features(k=3, s=0.8, deltas=d, normalize=False)
    res["nccl parts"] = sn.features(k=3, s=0.8, deltas=d, return_parts=True)[2][:, 0]
    res["nccl fused"] = sn.features(k=3, s=0.8, deltas=d, normalize=False)
    ref = res["single generic"]
    msg = f"[rank {rank}] seed {seed}: " + "; ".join(f"{k} {(v.reshape(ref.shape) - ref).abs().max().item():.2e}" for k, v in res.items())
    print(msg, flush=True)
sw.check_exchange()
dist.barrier()
os._exit(0)
