mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sell.py -x -q > gpurun_out/pytest_sell.log 2>&1; echo pytest-sell rc=$?
tail -3 gpurun_out/pytest_sell.log
for kb in 0 128 256 384 512 768; do
echo "== L2 prefetch $kb KB per CTA"
EGNN_SELL_L2_PREFETCH_KB=$kb timeout 600 python scripts/time_step.py reddit 3 2>&1 | grep -E "^order 2|spmv us|epilogue us|total us|graph"
done
