mkdir -p gpurun_out
S="--no-cpu-baseline --no-e2e --no-wide --no-ugca"
for cfg in "--steps 20 --warmup 5" "--steps 20 --warmup 5" "--steps 20 --warmup 50" "--steps 40 --warmup 5" "--steps 200 --warmup 10" "--steps 20 --warmup 5 --no-graph"; do
tag=$(echo $cfg | tr -d ' -')
timeout 300 python bench.py $S $cfg > gpurun_out/s20_$tag.log 2> gpurun_out/s20_$tag.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s20_$tag.log').read().strip().splitlines()[-1])
r=d['roofline']
print('$cfg', 'ms/step', round(d['ms_per_step'],5), 'launch_ms', round(r['avg_launch_ms'],5), 'share', round(r['kernel_share_of_step'],3), r['phase_us']['total'])
PY
done
