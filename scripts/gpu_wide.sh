mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest_gpu.log
for cfg in "arxiv 128 1" "reddit 64 1" "physics 8415 1" "arxiv 16 1" "reddit 8 2"; do
  set -- $cfg
  timeout 300 python bench.py --workload $1 --f $2 --scales $3 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/sweep_$1_f$2_s$3.log 2> gpurun_out/sweep_$1_f$2_s$3.err
  echo "$cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_$1_f$2_s$3.log').read().strip().splitlines()[-1])
    r=d['roofline']
    print('  ms/step %.4f value %.3e per-order %s frac %.3f share %.2f' % (d['ms_per_step'], d['value'], [round(x,4) for x in r['per_order_ms']], r['frac'], r['order_kernel_share_of_step']))
except Exception as e:
    print('  parse failed', e); print(open('gpurun_out/sweep_$1_f$2_s$3.err').read()[-800:])
PY
done
