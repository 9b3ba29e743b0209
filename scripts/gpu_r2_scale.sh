mkdir -p gpurun_out
for N in ${1:-8}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 10 --no-e2e > gpurun_out/bench_n${N}_reddit.log 2> gpurun_out/bench_n${N}_reddit.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_reddit.log').read().strip().splitlines()[-1])
    print('N=$N ms/step', d['ms_per_step'], 'check', d['check'], 'err', d['exchange_error'], 'ugca', (d.get('ugca') or {}).get('recompute_ms'))
    ph = d['run']['phase_us_rank0']
    print({k: v for k, v in ph.items() if k != 'per_cta_us_after_order_opened_min_mean_max'})
    for t in ph['per_cta_us_after_order_opened_min_mean_max']: print(t)
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_n${N}_reddit.err').read()[-1500:])
PY
done
