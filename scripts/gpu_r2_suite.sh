mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_surrogate.py -x -q > gpurun_out/pytest_sur.log 2>&1; echo pytest-surrogate rc=$?
grep -v "^epoch" gpurun_out/pytest_sur.log | tail -25
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?
grep -v "^epoch" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke.log
