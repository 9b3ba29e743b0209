# single-GPU sweep over the named shapes (kernel-only numbers; no CPU leg)
mkdir -p gpurun_out
for cfg in "arxiv 128 1" "arxiv 1 1" "reddit 64 1" "reddit 1 4" "physics 8415 1" "physics 1 1" "pubmed 1 1" "cora 1 1"; do
  set -- $cfg
  timeout 300 python bench.py --workload $1 --f $2 --scales $3 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/sweep_$1_f$2_s$3.log 2> gpurun_out/sweep_$1_f$2_s$3.err
  echo "$cfg rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sweep_$1_f$2_s$3.log').read().strip().splitlines()[-1])
    r=d['roofline']
    print('  ms/step %.4f value %.3e per-order %s frac %.3f share %.2f' % (d['ms_per_step'], d['value'], [round(x,4) for x in r['per_order_ms']], r['frac'], r['order_kernel_share_of_step']))
except Exception as e:
    print('  parse failed', e)
PY
done
