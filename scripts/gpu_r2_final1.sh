# final validation on 1 GPU: full suite (-s for the printed parity ratios), smoke, the driver's bench command, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_final.log 2>&1; echo pytest rc=$?
grep -v "^epoch\|^Early" gpurun_out/pytest_gpu_final.log | grep -E "passed|failed|error" | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo bench rc=$?; tail -2 gpurun_out/bench_final.err
timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final_ref.log 2> gpurun_out/bench_final_ref.err; echo ref rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/bench_final_ref.log').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value %.3e' % d['value'], 'frac', round(d['roofline']['frac'],3), 'spmv frac', round(d['roofline']['spmv_phase']['frac'],3), 'launches', d['gpu_launches'])
print('wide', [(w['workload'], w['f'], round(w['avg_launch_ms'],4), round(w['frac'],3)) for w in d['roofline_wide']])
print('ugca', d['ugca']['recompute_ms'], d['ugca']['e2e_ms'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e value %.3e' % d['e2e']['value'])
print('cpu_baseline', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline'].get('seconds'))
print('ref', r['value'], r['ms_per_step'], r['cpu_baseline']['kind'], r['config'] == d['config'])
print('e2e ratio', d['e2e']['value'] / r['value'], 'ratio', d['value'] / r['value'])
PY
