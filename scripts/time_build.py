"""Stage timing of the host-CSR entry on the Reddit shape (CUDA events + host clock)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth, _cabi

dev = torch.device("cuda", 0)
workload = sys.argv[1] if len(sys.argv) > 1 else "reddit"
rp, ci, n = synth.synth_csr(workload, self_loops=True, device=dev)
rp_h, ci_h = rp.cpu().pin_memory(), ci.cpu().pin_memory()
del rp, ci
torch.cuda.synchronize()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

for it in range(3):
    t0 = time.perf_counter(); e0 = ev()
    rpd = rp_h.to(dev, non_blocking=True); cid = ci_h.to(dev, non_blocking=True)
    e1 = ev()
    g = egnn.CsrGraph(rpd, cid, None, n)
    e2 = ev()
    plan = g.sell_plan()
    e3 = ev()
    feats = egnn.graph_wavelet_features(g)
    e4 = ev()
    out = feats.cpu()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"iter {it}: wall {1e3*(t1-t0):.2f} ms | h2d {e0.elapsed_time(e1):.2f} prep {e1.elapsed_time(e2):.2f} "
          f"plan {e2.elapsed_time(e3):.2f} orders {e3.elapsed_time(e4):.2f} | slices {plan.n_slices} vrows {plan.n_vrows} "
          f"entries {plan.n_entries} rowv {plan.n_rowv}", flush=True)

# finer: plan build pieces
lib = _cabi.load()
g = egnn.CsrGraph(rp_h.to(dev), ci_h.to(dev), None, n)
nb, cb, lmax = C.c_int32(0), C.c_int32(0), C.c_int32(0)
lib.egnn_sell_geometry(n, g.nnz, C.byref(nb), C.byref(cb), C.byref(lmax))
from efficient_gnn_b200.graph import build_sell_plan
torch.cuda.synchronize()
for it in range(2):
    t0 = time.perf_counter()
    p = build_sell_plan(g.rowptr, g.colidx, n, n, 0, (nb.value, cb.value, lmax.value))
    torch.cuda.synchronize()
    print(f"build_sell_plan wall {1e3*(time.perf_counter()-t0):.2f} ms", flush=True)
