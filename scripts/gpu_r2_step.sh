mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sell.py -x -q > gpurun_out/pytest_sell.log 2>&1; echo pytest-sell rc=$?
tail -3 gpurun_out/pytest_sell.log
timeout 600 python scripts/time_step.py reddit 3 > gpurun_out/time_step.log 2>&1; echo time rc=$?; tail -32 gpurun_out/time_step.log
