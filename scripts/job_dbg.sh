set -u
timeout 900 python -m pytest tests -m gpu -q -x -k "hybrid or mega_tcgen05" 2>&1 | tail -3
python bench.py --variant hybrid --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/bench_r2_hybrid.json 2> gpurun_out/bench_r2_hybrid.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r2_hybrid.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['breakdown_ms'], d['roofline']['kernel'][:80])
"
