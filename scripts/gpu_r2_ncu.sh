# round 2 captures: launch list + full capture of the step kernel (default bench), full captures of the wide kernel on the
# arxiv shape (dynamic row schedule) and of the plan-free blocked kernel (first use of the Reddit-shape graph)
mkdir -p gpurun_out
S="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-wide --no-ugca"
python bench.py $S > gpurun_out/plain_r2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sell_step -s 3 -c 2 -f -o gpurun_out/prof_step_r2 python bench.py $S > gpurun_out/ncu_step.log 2>&1
echo ncu-step rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py $S > gpurun_out/ncu_launches.log 2>&1
echo ncu-launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:blocked_spmv -c 2 -f -o gpurun_out/prof_blocked_r2 python bench.py $S > gpurun_out/ncu_blocked.log 2>&1
echo ncu-blocked rc=$?
A="--workload arxiv --f 128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $A > gpurun_out/plain_arxiv_r2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cheb_wide -s 6 -c 2 -f -o gpurun_out/prof_wide_arxiv_r2b python bench.py $A > gpurun_out/ncu_wide.log 2>&1
echo ncu-wide rc=$?
