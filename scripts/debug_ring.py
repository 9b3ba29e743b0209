import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASE = r'''
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
f, k, norm, shape = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
rp, ci, n = synth.synth_csr(shape, self_loops=True)
g = egnn.CsrGraph(rp.cuda(), ci.cuda(), None, n)
x0 = torch.randn(n, f, device="cuda")
out = egnn.graph_wavelet_features(g, k=k, X0=x0, normalize=bool(norm))
torch.cuda.synchronize()
print("ok", f, k, norm, shape, float(out.abs().sum()))
'''
for args in [(32,1,0,"cora"),(8,1,0,"cora")]:
    r = subprocess.run([sys.executable, "-c", CASE, *map(str,args)], capture_output=True, text=True, env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
    tail = (r.stdout.strip().splitlines() or [""])[-1]
    err = r.stderr.strip().splitlines()[-2:]
    print(args, "rc", r.returncode, tail, err)
