mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
grep -v "^epoch" gpurun_out/pytest_gpu.log | tail -6
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo bench rc=$?; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
r=d['roofline']
print('ms/step', d['ms_per_step'], 'launch ms', r['avg_launch_ms'], 'frac', r['frac'], 'frac_contract', r['frac_contract'])
print('phases', r['phase_us'])
print('spmv_phase', r['spmv_phase'])
for w in d['roofline_wide'] or []: print(w['workload'], w['f'], w['per_order_ms'], w['frac'], w['b_gather_achieved'])
print('ugca', d['ugca'])
print('e2e', d['e2e'])
print('cpu', d['cpu_baseline'])
PY
