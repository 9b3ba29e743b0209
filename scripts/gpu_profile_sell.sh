mkdir -p gpurun_out
S="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $S > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sell_spmv -s 9 -c 3 -f -o gpurun_out/prof_sell_v5 python bench.py $S > gpurun_out/ncu1.log 2>&1
echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v5.csv python bench.py $S > gpurun_out/ncu2.log 2>&1
echo ncu2 rc=$?
A="--workload arxiv --f 128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $A > gpurun_out/plain_arxiv.log 2>&1 &&
