"""Eager calls vs a replayed CUDA graph (WaveletSession) on the default workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
dev = torch.device("cuda", 0)
for wl in ("reddit", "arxiv", "pubmed", "cora"):
    rp, ci, n = synth.synth_csr(wl, self_loops=True, device=dev)
    g = egnn.CsrGraph(rp, ci, None, n)
    for _ in range(5): egnn.graph_wavelet_features(g)
    ses = egnn.WaveletSession(g, k=3, s=0.8)
    for name, fn in (("eager", lambda: egnn.graph_wavelet_features(g)), ("graph", lambda: ses())):
        for _ in range(10): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(200): fn()
        b.record(); torch.cuda.synchronize()
        print(f"{wl} {name}: {a.elapsed_time(b) / 200 * 1e3:.1f} us/step", flush=True)
