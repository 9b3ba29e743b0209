mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/e2e_n2.log 2> gpurun_out/e2e_n2.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/e2e_n2.log').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'check', d['check']['max_abs_diff_vs_single_gpu'], d['check'].get('ugca_max_abs_diff_vs_single_gpu'), 'err', d['exchange_error'], 'ugca', d['ugca']['recompute_ms'], 'e2e ms', d['e2e']['ms_per_step'])
PY
