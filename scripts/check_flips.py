"""Edge flips at Reddit scale: SELL step kernel vs generic kernel vs rebuilt graph (single GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
from bench import make_flips
dev = torch.device("cuda", 0)
rp, ci, n = synth.synth_csr("reddit", self_loops=True, device=dev)
g = egnn.CsrGraph(rp, ci, None, n)
for seed in (100, 101, 102):
    d = make_flips(n, 5, seed)
    print("flips", d)
    a = egnn.graph_wavelet_features(g, k=3, s=0.8, deltas=d, normalize=False, _use_sell=True)
    b = egnn.graph_wavelet_features(g, k=3, s=0.8, deltas=d, normalize=False, _use_sell=False)
    # rebuilt graph
    rows = torch.repeat_interleave(torch.arange(n, device=dev), (g.rowptr[1:] - g.rowptr[:-1]).long())
    keys = rows * n + g.colidx.long()
    add = torch.tensor([r * n + c for r, c in zip(d[0], d[1])], device=dev, dtype=torch.long)
    allk = torch.cat([keys, add])
    g2 = egnn.CsrGraph.from_edge_index(torch.stack([allk // n, allk % n]), n)
    c = egnn.graph_wavelet_features(g2, k=3, s=0.8, normalize=False)
    print("weighted rebuilt graph:", g2.vals is not None)
    for name, x, y in (("sell vs generic", a, b), ("sell vs rebuilt", a, c), ("generic vs rebuilt", b, c)):
        diff = (x - y).abs()
        i = int(diff.argmax())
        print(f"  {name}: max abs diff {diff.max().item():.3e} at row {i} (values {x.view(-1)[i].item():.6f} {y.view(-1)[i].item():.6f}); target {d[0][0]}")
