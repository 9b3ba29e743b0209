mkdir -p gpurun_out
A="--workload arxiv --f 128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py $A > gpurun_out/plain_arxiv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cheb_ring -s 10 -c 1 -f -o gpurun_out/prof_ring1_arxiv python bench.py $A > gpurun_out/ncu_arxiv.log 2>&1
echo arxiv rc=$?
