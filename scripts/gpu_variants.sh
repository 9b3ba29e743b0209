mkdir -p gpurun_out
python scripts/time_build.py 2>&1 | tee gpurun_out/time_build.log
B="--steps 50 --warmup 5 --no-cpu-baseline --no-e2e"
for st in 0.6180339887 0.0; do for v in 0 1 2 3 4; do
  EGNN_SELL_STRIDE=$st EGNN_SELL_VARIANT=$v python bench.py $B > gpurun_out/var.log 2>gpurun_out/var.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/var.log').read().strip().splitlines()[-1])
print('stride $st variant $v: ms/step %.4f per-order %s' % (d['ms_per_step'], [round(x,4) for x in d['roofline']['per_order_ms']]))
PY
done; done
