# final validation on 1 GPU: full suite (-s for the printed parity ratios), smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_final.log 2>&1; echo pytest rc=$?
grep -v "^epoch\|^Early" gpurun_out/pytest_gpu_final.log | grep -E "passed|failed|error" | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo bench rc=$?; tail -2 gpurun_out/bench_final.err
timeout 900 python bench.py > gpurun_out/bench_final_default.log 2> gpurun_out/bench_final_default.err; echo bench-default rc=$?
timeout 1500 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_final_ref.log 2> gpurun_out/bench_final_ref.err; echo ref rc=$?
