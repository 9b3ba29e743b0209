mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sell.py tests/test_gpu_wats.py tests/test_gpu_properties.py tests/test_gpu_sharded.py -q -x > gpurun_out/pytest_ugca.log 2>&1; echo pytest rc=$?
grep -v "^epoch\|^Early" gpurun_out/pytest_ugca.log | tail -2
timeout 300 python bench.py --steps 50 --warmup 5 --flips 5 --no-cpu-baseline --no-wide --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flips mode ms/step', d['ms_per_step'], d['roofline']['phase_us'])"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-wide --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', d['ms_per_step'], 'ugca', d['ugca']['recompute_ms'], d['ugca']['e2e_ms'])"
