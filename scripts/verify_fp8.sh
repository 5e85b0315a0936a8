#!/bin/bash
# Short check of the FP8 mode under gpurun: its GPU tests, the batch-1 bench line with the fp8_batch1 leg, then the rest of the suite.
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -q -k "fp8" -s > $O/r2_gputest_fp8b.log 2>&1; echo "pytest fp8 rc=$?"; grep -E "fp8 vs bf16|passed|failed|Error" $O/r2_gputest_fp8b.log | tail -6
timeout 200 python bench.py --no-batch64 --no-hybrid --no-ref-gpu --no-cpu-baseline > $O/bench_r2_fp8b.json 2> $O/bench_r2_fp8b.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_fp8b.json').read().strip().splitlines()[-1])
print('bs1', round(d['value'],2), round(d['roofline']['us_per_launch'],1)); print('fp8', json.dumps(d.get('fp8_batch1')))
PY
timeout 300 python -m pytest tests -m gpu -q -k "not fp8" > $O/r2_gputest_final_b.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2_gputest_final_b.log
