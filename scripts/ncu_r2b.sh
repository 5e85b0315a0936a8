#!/bin/bash
# Round-2 final captures (run under gpurun): default bench record, launch list of a short batch-1 pass, ncu --set full of the
# persistent decode step (transformer and hybrid stack) and of decode_tc_kernel at a mid-utterance context.  Every ncu run follows
# a plain run of the same command that exited 0.  The .ncu-rep files stay on the box: raw / details pages are exported as text.
set -u
O=gpurun_out
P=/tmp/zb_prof; mkdir -p $P
python bench.py > $O/bench_r2_final.json 2> $O/bench_r2_final.err || { echo "bench failed"; tail -5 $O/bench_r2_final.err; exit 1; }
S="python bench.py --steps 1 --warmup 1 --frames 24 --no-cpu-baseline --no-ref-gpu --no-batch64 --no-hybrid"
$S > $O/ncu_r2b_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file $O/r2_launches_b1.csv $S > $O/ncu_r2b_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_step_kernel -s 60 -c 1 -f -o $P/decode_step $S > $O/ncu_r2b_a.log 2>&1
H="python bench.py --variant hybrid --steps 1 --warmup 1 --frames 24 --no-cpu-baseline --no-ref-gpu"
$H > $O/ncu_r2b_hplain.log 2>&1 && ncu --set full --clock-control none -k regex:decode_step_kernel -s 60 -c 1 -f -o $P/decode_step_hybrid $H > $O/ncu_r2b_b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r2_launches_hybrid_b1.csv $H > $O/ncu_r2b_hl.log 2>&1
T="python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline --no-ref-gpu"
ncu --set full --clock-control none -k regex:decode_tc -s 430 -c 1 -f -o $P/decode_tc_mid $T > $O/ncu_r2b_c.log 2>&1
for n in decode_step decode_step_hybrid decode_tc_mid; do
  [ -f $P/$n.ncu-rep ] || continue
  ncu -i $P/$n.ncu-rep --page raw --csv > $O/r2b_ncu_${n}_raw.csv 2>/dev/null
  ncu -i $P/$n.ncu-rep --page details > $O/r2b_ncu_${n}_details.txt 2>/dev/null
done
ls -la $P $O/r2b_ncu_* | tail -12
