# cost-balanced epilogue ranges: narrow-path tests, default bench (regression check), then degree-sorted N = 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sell.py tests/test_gpu_full_size.py tests/test_gpu_properties.py -q -x > gpurun_out/pytest_epi.log 2>&1; echo pytest rc=$?
tail -3 gpurun_out/pytest_epi.log
S="--steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-wide"
timeout 600 python bench.py $S > gpurun_out/epi_default.log 2> gpurun_out/epi_default.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/epi_default.log').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'ugca', d['ugca']['recompute_ms'], d['roofline']['phase_us'])
PY
