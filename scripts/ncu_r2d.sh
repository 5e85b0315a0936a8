#!/bin/bash
# FP8 mode of the persistent decode step (ZB_FP8=1), under gpurun: bench line with the fp8_batch1 leg (tolerance figures on identical
# histories), per-phase timeline of CTA 0 in both modes, ncu --set full of one FP8 launch (after a plain run of the same command).
set -u
O=gpurun_out
P=/tmp/zb_prof; mkdir -p $P $O
timeout 300 python bench.py --no-batch64 --no-hybrid --no-ref-gpu --no-cpu-baseline > $O/bench_r2_fp8.json 2> $O/bench_r2_fp8.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_fp8.json').read().strip().splitlines()[-1])
print('bs1', round(d['value'],2), round(d['roofline']['us_per_launch'],1)); print('fp8', json.dumps(d.get('fp8_batch1')))
PY
ZB_FP8=0 timeout 200 python scripts/timeline_mega.py > $O/r2d_timeline_bf16.txt 2>&1; echo "tl bf16 rc=$?"; tail -6 $O/r2d_timeline_bf16.txt
ZB_FP8=1 timeout 200 python scripts/timeline_mega.py > $O/r2d_timeline_fp8.txt 2>&1; echo "tl fp8 rc=$?"; tail -6 $O/r2d_timeline_fp8.txt
ZB_FP8=1 ZB_TL_N=200 timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 150 -c 1 -f -o $P/decode_step_fp8 python scripts/timeline_mega.py > $O/ncu_r2d.log 2>&1; echo "ncu rc=$?"
if [ -f $P/decode_step_fp8.ncu-rep ]; then
  ncu -i $P/decode_step_fp8.ncu-rep --page raw --csv > $O/r2d_ncu_decode_step_fp8_raw.csv 2>/dev/null
  ncu -i $P/decode_step_fp8.ncu-rep --page details > $O/r2d_ncu_decode_step_fp8_details.txt 2>/dev/null
  grep -E "decode_step_kernel|Duration|DRAM Throughput|dram__bytes_read.sum |Registers Per|Achieved Occupancy" $O/r2d_ncu_decode_step_fp8_details.txt | head -12
fi
