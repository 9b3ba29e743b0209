# last validation of the round on 1 GPU: the whole GPU suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_last.log 2>&1; echo pytest rc=$?
grep -v "^epoch\|^Early" gpurun_out/pytest_gpu_last.log | grep -E "passed|failed|error|^E " | tail -12
