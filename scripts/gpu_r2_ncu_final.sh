# end-of-round captures on the final tree: launch list + full capture of the step kernel (default bench arguments)
mkdir -p gpurun_out
S="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-wide --no-ugca"
python bench.py $S > gpurun_out/plain_r2f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sell_step -s 3 -c 2 -f -o gpurun_out/prof_step_r2f python bench.py $S > gpurun_out/ncu_step_f.log 2>&1
echo ncu-step rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2f.csv python bench.py $S > gpurun_out/ncu_launches_f.log 2>&1
echo ncu-launches rc=$?
