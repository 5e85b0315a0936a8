set -u
O=gpurun_out
( time python bench.py > $O/bench_r2_b.json 2> $O/bench_r2_b.err ) 2>&1 | tail -3
tail -c 600 $O/bench_r2_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_b.json').read().strip().splitlines()[-1])
print('bs1', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['breakdown_ms'], d['ttfa']['p50_ms'])
b=d.get('batch64'); 
if b: print('bs64', b['value'], b['e2e']['value'], b['roofline']['frac'], b['roofline']['us_per_launch'], b['breakdown_ms'])
h=d.get('hybrid')
if h: print('hybrid', h['value'], h['roofline']['frac'], h['roofline']['us_per_launch'], h['breakdown_ms'])
print('ref_gpu', d.get('reference_gpu_eager')); print('cpu', d.get('cpu_baseline'))
PY
