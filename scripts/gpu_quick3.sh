mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sell.py tests/test_gpu_sharded.py tests/test_gpu_peer.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench.log 2> gpurun_out/bench.err; echo bench rc=$?; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['per_order_ms'], d['roofline']['frac'], d['roofline']['order_kernel_share_of_step'])
PY
for w in 1 8; do python scripts/profile_shard.py $w 50 2>&1 | tail -2; done
python scripts/profile_shard.py 8 3 > gpurun_out/plain_shard.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_shard8.csv python scripts/profile_shard.py 8 3 > gpurun_out/ncu_shard.log 2>&1
echo ncu rc=$?
