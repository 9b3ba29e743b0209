N=${1:-8}
mkdir -p gpurun_out
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $1 --f $2 --steps ${3:-100} --warmup 5 --check --no-cpu-baseline $4 > gpurun_out/bench_n${N}_$1_f$2.log 2> gpurun_out/bench_n${N}_$1_f$2.err
  echo "bench $1 f=$2 rc=$?"
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/bench_n${N}_$1_f$2.log') if l.startswith('{')][-1]
    print('  N=%d %s/%s: ms/step %.4f value %.3e check %s err %s e2e_ms %s' % (d['n_gpus'], d['config'].get('path'), d['config'].get('exchange'), d['ms_per_step'], d['value'], d.get('check'), d.get('exchange_error'), d['e2e'] and round(d['e2e']['ms_per_step'],2)))
except Exception as e:
    print('  parse failed', e); print(open('gpurun_out/bench_n${N}_$1_f$2.err').read()[-1500:])
PY
}
run reddit 1 3000 --no-e2e
run reddit 64 20
run arxiv 128 50
