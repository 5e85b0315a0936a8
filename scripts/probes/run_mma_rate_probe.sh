#!/bin/bash
# tcgen05.mma issue cost per instruction (scripts/probes/mma_rate_probe.cu): lane-0 style issue (3) against the elect.sync pattern (4)
P=scripts/probes/mma_rate_probe
for args in "128 16 512 1 3" "128 16 512 1 4" "64 16 512 1 4" "64 64 512 1 4" "64 128 512 1 4" "64 192 512 1 4" "64 224 512 1 4" "64 256 512 1 4" "128 128 512 1 4" "128 256 512 1 4" "64 256 8 1 4" "64 256 32 1 4"; do
  $P $args
done
