#!/bin/bash
# Round-end check on a GPU box (under gpurun): the GPU test suite (the opt-in FP8 mode in a process of its own, so that a fault
# there cannot take the other tests with it), smoke(), the default bench line, gemm_tc against torch.matmul.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "not fp8" > $O/r2_gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_gputest_final.log
timeout 400 python -m pytest tests -m gpu -q -k "fp8" -s > $O/r2_gputest_fp8.log 2>&1; FP8RC=$?; echo "pytest fp8 rc=$FP8RC"; grep -E "fp8 vs bf16|passed|failed|Error|error" $O/r2_gputest_fp8.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
FP8FLAG=""; if [ $FP8RC -ne 0 ]; then FP8FLAG="--no-fp8"; fi
timeout 900 python bench.py $FP8FLAG > $O/bench_r2_final3.json 2> $O/bench_r2_final3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final3.json').read().strip().splitlines()[-1])
print('bs1', round(d['value'],2), round(d['e2e']['value'],2), round(d['roofline']['frac'],3), round(d['roofline']['us_per_launch'],1), d['breakdown_ms'], round(d['ttfa']['p50_ms'],1), d['roofline']['traffic'])
b=d['batch64']; print('bs64', round(b['value'],1), round(b['e2e']['value'],1), round(b['roofline']['frac'],3), round(b['roofline']['us_per_launch'],1), b['breakdown_ms'], b['roofline']['traffic'])
h=d['hybrid_batch1']; print('hyb', round(h['value'],2), round(h['roofline']['frac'],3), round(h['roofline']['us_per_launch'],1), h['breakdown_ms'], h['roofline']['traffic'])
print('fp8', json.dumps(d.get('fp8_batch1')))
PY
timeout 240 python scripts/time_gemm_tc.py > $O/gemm_vs_matmul.txt 2>&1; echo "gemm rc=$?"; cat $O/gemm_vs_matmul.txt
