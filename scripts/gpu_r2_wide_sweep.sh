mkdir -p gpurun_out
for lib in libegnn_b200.so libegnn_b200_u6b3.so libegnn_b200_u8b3.so libegnn_b200_u8b2.so libegnn_b200_u12b2.so; do
for cfg in "--workload arxiv --f 128" "--workload reddit --f 64"; do
echo "== $lib $cfg"
EGNN_LIB_PATH=$PWD/efficient-gnn_b200/lib/$lib timeout 600 python bench.py $cfg --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step', round(d['ms_per_step'],4), 'per order', [round(x,4) for x in r['per_order_ms']], 'frac', round(r['frac'],4))"
done; done
