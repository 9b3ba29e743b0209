mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_ugca.log 2>&1; echo pytest rc=$?
grep -v "^epoch\|^Early" gpurun_out/pytest_ugca.log | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide > gpurun_out/bench_ugca.log 2>gpurun_out/bench_ugca.err; echo bench rc=$?; tail -2 gpurun_out/bench_ugca.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ugca.log').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'ugca', d['ugca']['recompute_ms'], d['ugca']['e2e_ms'], 'e2e', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
PY
timeout 300 python bench.py --steps 50 --warmup 5 --flips 5 --no-cpu-baseline --no-wide --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flips mode ms/step', d['ms_per_step'], d['roofline']['phase_us'])"
