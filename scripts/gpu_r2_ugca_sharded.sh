# sharded UGCA recompute with in-place node patches: the 2-rank flip tests, then the N = 2 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py -q -x -k "True or stale or light" > gpurun_out/pytest_peer_flips.log 2>&1; echo pytest rc=$?
tail -4 gpurun_out/pytest_peer_flips.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 10 --no-e2e > gpurun_out/ugca_n2.log 2> gpurun_out/ugca_n2.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/ugca_n2.log').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'check', d['check'], 'err', d['exchange_error'], 'ugca', d['ugca']['recompute_ms'])
PY
