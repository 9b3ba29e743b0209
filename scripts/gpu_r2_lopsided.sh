# the lopsided-shard regression test against a library built from the previous sell_step.cuh (scratch_lib/, if present:
# it failed there with 4.7e-2 on the heavy rank) and the current one,
# then the 2-rank suite and the F = 1 bench at N = 2
mkdir -p gpurun_out
if [ -f scratch_lib/libegnn_b200_prev.so ]; then
EGNN_LIB_PATH=$PWD/scratch_lib/libegnn_b200_prev.so timeout 600 python -m pytest tests/test_gpu_peer.py -q -k light_shard > gpurun_out/lopsided_prev.log 2>&1; echo prev rc=$?
grep -E "passed|failed|assert .*worst|^E " gpurun_out/lopsided_prev.log | head -5
fi
timeout 1500 python -m pytest tests/test_gpu_peer.py -q > gpurun_out/pytest_peer_final.log 2>&1; echo pytest-peer rc=$?
tail -3 gpurun_out/pytest_peer_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/lop_n2.log 2> gpurun_out/lop_n2.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/lop_n2.log').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'check', d['check'], 'err', d['exchange_error'], 'ugca', d['ugca']['recompute_ms'])
print({k: v for k, v in d['run']['phase_us_rank0'].items() if k != 'per_cta_us_after_order_opened_min_mean_max'})
PY
