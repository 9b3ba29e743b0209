# final multi-GPU campaign: for N in "$@": (N=2: the 2-rank tests), then the three sharded configs of DESIGN.md section 6
mkdir -p gpurun_out
for N in "$@"; do
if [ "$N" = "2" ]; then
timeout 1500 python -m pytest tests/test_gpu_peer.py -q > gpurun_out/pytest_peer_final.log 2>&1; echo pytest-peer rc=$?
tail -2 gpurun_out/pytest_peer_final.log
fi
for cfg in "--workload reddit" "--workload reddit --f 64" "--workload arxiv --f 128"; do
tag=$(echo $cfg | tr -d ' -')
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 10 $cfg > gpurun_out/final_n${N}_$tag.log 2> gpurun_out/final_n${N}_$tag.err; echo "bench N=$N $cfg rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/final_n${N}_$tag.log').read().strip().splitlines()[-1])
    print('  ms/step', round(d['ms_per_step'],5), 'value', '%.3e' % d['value'], 'check', d['check'], 'err', d['exchange_error'], 'ugca', (d.get('ugca') or {}).get('recompute_ms'), 'e2e ms', (d.get('e2e') or {}).get('ms_per_step'), d['run']['path'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/final_n${N}_$tag.err').read()[-1500:])
PY
done
done
