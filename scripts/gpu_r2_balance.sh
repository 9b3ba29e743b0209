# equal-rows shards of a degree-sorted numbering, without and with sharded.BalancedOrder (N = $1, default 2)
mkdir -p gpurun_out
N=${1:-2}
for cfg in ${CFGS:-"--workload reddit --node-order degree" "--workload reddit --node-order degree --balance" "--workload reddit --f 64 --node-order degree" "--workload reddit --f 64 --node-order degree --balance"}; do
tag=$(echo $cfg | tr -d ' -')
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 5 --no-e2e $cfg > gpurun_out/bal_n${N}_$tag.log 2> gpurun_out/bal_n${N}_$tag.err; echo "bench N=$N $cfg rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bal_n${N}_$tag.log').read().strip().splitlines()[-1])
    print('  ms/step', round(d['ms_per_step'],5), 'check', d['check']['max_abs_diff_vs_single_gpu'], d['check'].get('ugca_max_abs_diff_vs_single_gpu'), 'err', d['exchange_error'], d['run']['path'], d['run']['ordering'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bal_n${N}_$tag.err').read()[-1500:])
PY
done
