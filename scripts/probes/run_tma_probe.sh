#!/bin/bash
# HBM streaming rate by access pattern (see tma_tensor_probe.cu).  Output: gpurun_out/tma_probe.log
P=scripts/probes/tma_tensor_probe
OUT=gpurun_out/tma_probe.log
mkdir -p gpurun_out; : > $OUT
run() { timeout 60 $P "$@" >> $OUT 2>&1; }
# mode RB kps stages K total_MB mmaN promo
run 0 128 1 12 2048 2048 0
run 0 128 2 6 2048 2048 0
run 1 128 1 12 2048 2048 0
run 1 128 1 12 2048 2048 0 0
run 1 128 1 12 2048 2048 0 1
run 1 128 1 12 2048 2048 0 3
run 1 128 2 6 2048 2048 0
run 1 128 4 3 2048 2048 0
run 1 56 1 12 2048 2048 0
run 1 56 2 12 2048 2048 0
run 1 16 8 12 2048 2048 0
run 1 128 1 12 8192 2048 0
run 1 128 2 6 8192 2048 0
run 3 128 2 6 2048 2048 0
run 3 128 4 3 2048 2048 0
run 3 56 4 6 2048 2048 0
run 2 128 2 6 2048 2048 0
run 2 128 4 3 2048 2048 0
run 2 16 8 12 2048 2048 0
run 2 16 16 6 2048 2048 0
run 2 56 4 6 2048 2048 0
# with the tensor-core consumer
run 1 128 1 12 2048 2048 16
run 1 128 1 12 2048 2048 128
run 2 128 2 6 2048 2048 16
run 2 128 2 6 2048 2048 128
run 0 128 2 6 2048 2048 16
run 2 16 8 12 2048 2048 16
# fewer CTAs pulling
run 0 128 2 6 2048 2048 0 2 128
run 0 128 2 6 2048 2048 0 2 74
cat $OUT
