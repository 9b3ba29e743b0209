mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sell.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q > gpurun_out/pytest_blocked.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_blocked.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-ugca > gpurun_out/bench_blocked.log 2>gpurun_out/bench_blocked.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_blocked.log').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'])
PY
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import efficient_gnn_b200 as egnn
from efficient_gnn_b200 import synth
rp, ci, n = synth.synth_csr("reddit", self_loops=True, device="cuda")
g = egnn.CsrGraph(rp, ci, None, n)
for mode in ("blocked", False):
    for _ in range(3): egnn.graph_wavelet_features(g, _use_sell=mode)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): egnn.graph_wavelet_features(g, _use_sell=mode)
    b.record(); torch.cuda.synchronize()
    print(mode, a.elapsed_time(b) / 20, "ms per 3-order pass")
PY
