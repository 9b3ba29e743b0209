mkdir -p gpurun_out
timeout 600 python scripts/time_step.py reddit 3 2>&1 | grep -E "^order|spmv us|epilogue us|total us|graph|eager"
