mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py -q -x -k "18 or 19" > gpurun_out/pytest_peer_k.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/pytest_peer_k.log
