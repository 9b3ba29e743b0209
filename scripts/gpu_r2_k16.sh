mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sell.py -q -x -k "more_orders or degree_sorted" > gpurun_out/pytest_k16.log 2>&1; echo pytest rc=$?
tail -15 gpurun_out/pytest_k16.log
