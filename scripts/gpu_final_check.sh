mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print('value %.4e ms/step %.4f frac %.3f e2e %.4e (%.2f ms) cpu %.3e (%.1fs) launches %d clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['cpu_baseline']['seconds'], d['gpu_launches'], d['clocks']))
PY
