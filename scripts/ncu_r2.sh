#!/bin/bash
# Round-2 ncu captures (run under gpurun): decode_tc_kernel at batch 64, gemm_tc_kernel + attn_prefill_kernel (prefill),
# conv_tc_kernel (DAC), and the launch list of a short batch-64 pass.  Every ncu run follows a plain run of the same
# command that exited 0.  The .ncu-rep files stay on the box (gpurun_out/ is limited to 64 MiB): the raw / details pages
# are exported as text.
set -u
O=gpurun_out
P=/tmp/zb_prof; mkdir -p $P
CMD="python bench.py --batch 64 --steps 1 --warmup 1 --frames 24 --no-cpu-baseline --no-ref-gpu"
$CMD > $O/ncu_r2_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/ncu_r2_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r2_launches_b64.csv $CMD > $O/ncu_r2_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_tc -s 40 -c 2 -f -o $P/decode_tc $CMD > $O/ncu_r2_a.log 2>&1
ncu --set full --clock-control none -k "regex:gemm_tc|attn_prefill" -s 30 -c 8 -f -o $P/prefill $CMD > $O/ncu_r2_b.log 2>&1
DCMD="python scripts/dac_times.py --batch 4 --frames 431 --reps 1"
$DCMD > $O/ncu_r2_dac_plain.log 2>&1 && ncu --set full --clock-control none -k regex:conv_tc -s 29 -c 29 -f -o $P/conv_tc $DCMD > $O/ncu_r2_c.log 2>&1
for n in decode_tc prefill conv_tc; do
  [ -f $P/$n.ncu-rep ] || continue
  ncu -i $P/$n.ncu-rep --page raw --csv > $O/r2_ncu_${n}_raw.csv 2>/dev/null
  ncu -i $P/$n.ncu-rep --page details > $O/r2_ncu_${n}_details.txt 2>/dev/null
done
ncu -i $P/decode_tc.ncu-rep --page source --csv > $O/r2_ncu_decode_tc_source.csv 2>/dev/null
ls -la $P $O/r2_ncu_* | tail -12
