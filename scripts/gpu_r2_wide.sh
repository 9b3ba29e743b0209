mkdir -p gpurun_out
for cfg in "--workload arxiv --f 128" "--workload reddit --f 64" "--workload physics --f 8415"; do
echo "== $cfg"
timeout 600 python bench.py $cfg --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step', round(d['ms_per_step'],4), 'per order', [round(x,4) for x in r['per_order_ms']], 'frac', round(r['frac'],4), 'gather TB/s', round(r['b_gather_achieved']/1e3,2))"
done
