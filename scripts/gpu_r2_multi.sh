N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/pytest_peer.log 2>&1; echo pytest-peer rc=$?
tail -5 gpurun_out/pytest_peer.log
for i in 1 2; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/check_flips_sharded.py reddit 2>&1 | grep "rank 0\|Error\|error"; done
bash scripts/gpu_r2_scale.sh $N
