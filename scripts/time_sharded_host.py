"""Host time of one ShardedWavelet.features(deltas=...) call (enqueue only) against its device time:
tells whether the sharded UGCA recompute is bound by the Python side.  One GPU, world = 1 over NCCL."""
import cProfile
import os
import pstats
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from efficient_gnn_b200 import sharded, synth  # noqa: E402

os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", RANK="0", WORLD_SIZE="1")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
rp, ci, n = synth.synth_csr("reddit", self_loops=True, device=dev)
plain = sharded.ShardedWavelet(rp, ci, n, device=dev, peer_exchange=False)
peer = sharded.PeerExchange(plain.part.rows_per, 1, device=dev)
sw = sharded.ShardedWavelet(rp, ci, n, device=dev, peer_exchange=peer)
cands = [bench.make_flips(n, 5, 100 + i) for i in range(32)]
for mode, kw in (("unperturbed", {}), ("flips", None)):
    for i in range(5):
        sw.features(k=3, s=0.8, **({"deltas": cands[i]} if kw is None else kw))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(200):
        sw.features(k=3, s=0.8, **({"deltas": cands[i % 32]} if kw is None else kw))
    t_host = (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 200
    print(f"{mode}: host enqueue {t_host * 1e6:.1f} us per call, with the device {t_all * 1e6:.1f} us")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(300):
    sw.features(k=3, s=0.8, deltas=cands[i % 32])
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
dist.destroy_process_group()
