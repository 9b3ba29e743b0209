"""Where does CsrGraph construction spend its time?  Host-clock per call, device synchronised between calls."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth, _cabi
dev = torch.device("cuda", 0)
rp, ci, n = synth.synth_csr("reddit", self_loops=True, device=dev)
lib = _cabi.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {label}: {1e3*(time.perf_counter()-t0):.3f} ms", flush=True); return r
for it in range(3):
    print("iter", it)
    T("require_device", _cabi.require_device)
    dinv = T("empty x7", lambda: [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(5)] + [torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.float64, device=dev), torch.empty(1, dtype=torch.int32, device=dev)])
    d = dinv
    T("graph_prep", lambda: _cabi.check(lib.egnn_graph_prep(_cabi.ptr(rp), _cabi.ptr(ci), None, n, _cabi.ptr(d[0]), _cabi.ptr(d[5]), _cabi.ptr(d[1]), _cabi.ptr(d[2]), _cabi.ptr(d[3]), _cabi.ptr(d[4]), _cabi.ptr(d[6]), _cabi.ptr(d[7]), st)))
    T("graph_prep no-flag", lambda: _cabi.check(lib.egnn_graph_prep(_cabi.ptr(rp), _cabi.ptr(ci), None, n, _cabi.ptr(d[0]), _cabi.ptr(d[5]), _cabi.ptr(d[1]), _cabi.ptr(d[2]), _cabi.ptr(d[3]), _cabi.ptr(d[4]), _cabi.ptr(d[6]), None, st)))
    g = T("CsrGraph()", lambda: egnn.CsrGraph(rp, ci, None, n))
    T("sell_plan", lambda: g.sell_plan())
    T("features", lambda: egnn.graph_wavelet_features(g))
