mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_patch1.log 2>&1; echo pytest rc=$?
grep -v "^epoch\|^Early" gpurun_out/pytest_patch1.log | tail -5
S="--steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-wide"
timeout 600 python bench.py $S > gpurun_out/patch1_bench.log 2> gpurun_out/patch1_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/patch1_bench.log').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'ugca', d['ugca']['recompute_ms'], d['ugca']['e2e_ms'], d['roofline']['phase_us'])
PY
