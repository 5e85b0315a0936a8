#!/bin/bash
# Round-end check on a GPU box (under gpurun): the GPU test suite, smoke(), the default bench line and BASELINE.json configs[4].
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_gputest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/bench_r2_final2.json 2> $O/bench_r2_final2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final2.json').read().strip().splitlines()[-1])
print('bs1', round(d['value'],2), round(d['e2e']['value'],2), round(d['roofline']['frac'],3), round(d['roofline']['us_per_launch'],1), d['breakdown_ms'], round(d['ttfa']['p50_ms'],1), d['roofline']['traffic'])
b=d['batch64']; print('bs64', round(b['value'],1), round(b['e2e']['value'],1), round(b['roofline']['frac'],3), round(b['roofline']['us_per_launch'],1), b['breakdown_ms'], b['roofline']['traffic'])
h=d['hybrid_batch1']; print('hyb', round(h['value'],2), round(h['roofline']['frac'],3), round(h['roofline']['us_per_launch'],1), h['breakdown_ms'], h['roofline']['traffic'])
PY
python bench.py --variant hybrid --batch 32 --prefix-frames 258 --frames 2584 --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu > $O/bench_r2_cfg5b.json 2> $O/bench_r2_cfg5b.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_cfg5b.json').read().strip().splitlines()[-1])
print('cfg5', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), round(d['roofline']['us_per_launch'],1), d['breakdown_ms'])
PY
