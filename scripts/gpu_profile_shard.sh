mkdir -p gpurun_out
for w in 1 2 4 8; do python scripts/profile_shard.py $w 50 2>&1 | tail -3; done
python scripts/profile_shard.py 2 3 > gpurun_out/plain_shard.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_shard2.csv python scripts/profile_shard.py 2 3 > gpurun_out/ncu_shard.log 2>&1
echo ncu rc=$?
