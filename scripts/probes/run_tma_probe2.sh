#!/bin/bash
# TMA engine rate with L2-resident data (64 MB walked 16 times): tensor boxes of 128-byte rows vs 1-D bulk copies
P=scripts/probes/tma_tensor_probe
OUT=gpurun_out/tma_probe2.log
mkdir -p gpurun_out; : > $OUT
run() { timeout 60 $P "$@" >> $OUT 2>&1; }
# mode RB kps stages K total_MB mmaN promo grid loops
run 0 128 1 12 2048 74 0 2 148 16
run 0 128 2 6 2048 74 0 2 148 16
run 1 128 1 12 2048 74 0 2 148 16
run 1 128 2 6 2048 74 0 2 148 16
run 1 128 4 3 2048 74 0 2 148 16
run 3 128 4 3 2048 74 0 2 148 16
run 2 128 2 6 2048 74 0 2 148 16
run 2 128 4 3 2048 74 0 2 148 16
run 1 128 1 12 2048 74 128 2 148 16
run 0 128 1 12 2048 74 128 2 148 16
run 1 64 1 12 2048 74 0 2 148 16
run 1 128 1 12 2048 2048 0 2 74 1
run 1 128 2 6 2048 2048 0 2 74 1
run 0 128 1 12 2048 2048 0 2 74 1
cat $OUT
