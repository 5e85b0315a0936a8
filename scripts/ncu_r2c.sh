#!/bin/bash
# Round-2 DAC capture after the epilogue / halo changes (run under gpurun): ncu --set full of the 29 conv_tc_kernel launches of
# one decode (4 x 431 frames), after a plain run of the same command.
set -u
O=gpurun_out
P=/tmp/zb_prof; mkdir -p $P
DCMD="python scripts/dac_times.py --batch 4 --frames 431 --reps 1"
$DCMD > $O/ncu_r2c_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none -k regex:conv_tc -s 29 -c 29 -f -o $P/conv_tc2 $DCMD > $O/ncu_r2c.log 2>&1
ncu -i $P/conv_tc2.ncu-rep --page raw --csv > $O/r2c_ncu_conv_tc_raw.csv 2>/dev/null
ls -la $O/r2c_ncu_conv_tc_raw.csv; cat $O/ncu_r2c_plain.log | tail -1
