# usage: bash scripts/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > gpurun_out/pytest_peer.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_peer.log
for ex in peer nccl; do
  EGNN_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 5 --check > gpurun_out/bench_n${N}_$ex.log 2> gpurun_out/bench_n${N}_$ex.err
  echo "bench $ex rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/bench_n${N}_$ex.log') if l.startswith('{')][-1]
    print('  N=%d %s: ms/step %.4f value %.3e check %s err %s e2e %s' % (d['n_gpus'], d['config'].get('exchange'), d['ms_per_step'], d['value'], d.get('check'), d.get('exchange_error'), d['e2e'] and round(d['e2e']['ms_per_step'],2)))
except Exception as e:
    print('  parse failed', e); print(open('gpurun_out/bench_n${N}_$ex.err').read()[-1500:])
PY
done
