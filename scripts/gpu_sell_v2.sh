mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sell.py tests/test_gpu_properties.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sell.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_sell.log
B="--steps 100 --warmup 5 --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/bench.log 2> gpurun_out/bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['per_order_ms'], d['roofline']['frac'], d['roofline']['order_kernel_share_of_step'], d['e2e'])
PY
S="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $S > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 9 -c 3 -f -o gpurun_out/prof_sell_v2 python bench.py $S > gpurun_out/ncu1.log 2>&1
echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v2.csv python bench.py $S > gpurun_out/ncu2.log 2>&1
echo ncu2 rc=$?
