#!/bin/bash
# decode_tc.cu knobs at full size, batch 64 (gpurun_out/tc_sweep.log)
OUT=gpurun_out/tc_sweep.log; : > $OUT
run() { echo "== $*" >> $OUT; env "$@" timeout 300 python scripts/tc_check.py --dims full --batch ${BB:-64} --frames ${FR:-40} --cond-len 160 --time --no-compare ${TL:-} 2>&1 | grep -v "CUDAEvent.h" >> $OUT; }
timeout 300 python scripts/tc_check.py --dims tiny --batch 1 3 8 64 --frames 12 --oracle 2>&1 | cut -c1-330 >> $OUT
TL=--timeline FR=440 run ZB_TC_L2PF=1
FR=440 run ZB_TC_L2PF=0
FR=440 run ZB_TC_L2PF=1 ZB_TC_NA=8
FR=440 run ZB_TC_L2PF=1 ZB_TC_BSTAGES=2
cat $OUT
