"""Phase breakdown of the narrow-path step kernel (globaltimer stamps of CTA 0) + CUDA-graph step time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
from efficient_gnn_b200.graph import enable_phase_stamps, read_phase_stamps
dev = torch.device("cuda", 0)
wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rp, ci, n = synth.synth_csr(wl, self_loops=True, device=dev)
g = egnn.CsrGraph(rp, ci, None, n)
for _ in range(5): egnn.graph_wavelet_features(g, k=k, _use_sell=True)
plan = g.sell_plan()
print(f"{wl}: n={n} nnz={g.nnz} entries={plan.n_entries} slices={plan.n_slices} vrows={plan.n_vrows} rowv={plan.n_rowv} blocks={plan.n_blocks}x{plan.col_block}")
info = plan._keepalive["cta_info"].cpu().numpy()
print("CTAs per block:", np.bincount(info[:plan.n_cta][info[:plan.n_cta] >= 0]))
enable_phase_stamps(plan)
acc = []
for _ in range(20):
    egnn.graph_wavelet_features(g, k=k, _use_sell=True)
    torch.cuda.synchronize()
    acc.append(read_phase_stamps(plan, k, first_operand_in_kernel=False))
# per-CTA trace of one launch: [stage done, slices done, rows done] per order
enable_phase_stamps(plan, True, trace=True)
egnn.graph_wavelet_features(g, k=k, _use_sell=True)
torch.cuda.synchronize()
st = plan._keepalive["stamps"].cpu().numpy().astype(np.int64)
glob, per = st[:64], st[64:].reshape(plan.n_cta, 64)
blk = info[:plan.n_cta]
# global stamps: [0] start, [1] after prologue barrier, then per order: after barrier 1, after barrier 2 (or end)
for order in range(1, k + 1):
    rel0 = glob[0 + 2 * (order - 1)]                 # start of the order on CTA 0 (kernel start / its own rows of the previous order done)
    rel1 = glob[1 + 2 * (order - 1)]                 # release of barrier 1
    stage, done, rows = per[:, 3 * (order - 1)], per[:, 3 * (order - 1) + 1], per[:, 3 * (order - 1) + 2]
    f = lambda x: f"min {x.min() / 1e3:.1f} mean {x.mean() / 1e3:.1f} max {x.max() / 1e3:.1f}"
    print(f"order {order}: stage done after open: {f(stage - rel0)}; slices done after open: {f(done - rel0)}; "
          f"barrier1 release after last arrival: {(rel1 - done.max()) / 1e3:.1f}; rows done after release: {f(rows - rel1)}")
    for c in range(plan.n_blocks):
        m = blk == c
        print(f"   block {c}: {m.sum()} CTAs, slices done {f((done - rel0)[m])}")
enable_phase_stamps(plan, False)
med = lambda xs: float(np.median(xs))
print("prologue us", med([a["prologue"] for a in acc]))
print("spmv us", [med([a["spmv"][i] for a in acc]) for i in range(k)])
print("epilogue us", [med([a["epilogue"][i] for a in acc]) for i in range(k)])
print("total us", med([a["total"] for a in acc]))
ses = egnn.WaveletSession(g, k=k, s=0.8)
for name, fn in (("eager", lambda: egnn.graph_wavelet_features(g, k=k)), ("graph", lambda: ses())):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{wl} {name}: {a.elapsed_time(b) / 200 * 1e3:.1f} us/step", flush=True)
