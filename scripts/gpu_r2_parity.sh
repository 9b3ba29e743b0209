# round 2, call 1: the new full-size parity tests + the reference's own WATS class, then a fresh
# --set full capture of the SELL SpMV (baseline for the round's kernel work)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_reference_class.py -x -q -s > gpurun_out/pytest_r2_parity.log 2>&1; echo pytest rc=$?
grep -E "^\[|passed|failed|error|oracle operator" gpurun_out/pytest_r2_parity.log | tail -40
S="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $S > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 9 -c 2 -f -o gpurun_out/prof_sell_r2_base python bench.py $S > gpurun_out/ncu1.log 2>&1
echo ncu rc=$?
