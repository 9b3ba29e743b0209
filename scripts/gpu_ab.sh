# A/B of two library builds on one box: default bench, alternating, $1 rounds
mkdir -p gpurun_out
S="--steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-wide --no-check"
for i in $(seq 1 ${1:-2}); do
for v in eq cur; do
if [ $v = eq ]; then export EGNN_LIB_PATH=$PWD/scratch_lib/libegnn_b200_eq.so; else unset EGNN_LIB_PATH; fi
timeout 300 python bench.py $S > gpurun_out/ab_${v}_$i.log 2> gpurun_out/ab_${v}_$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/ab_${v}_$i.log').read().strip().splitlines()[-1])
print('$v', $i, 'ms/step', round(d['ms_per_step'],5), 'ugca', round(d['ugca']['recompute_ms'],5), d['roofline']['phase_us'])
PY
done
done
