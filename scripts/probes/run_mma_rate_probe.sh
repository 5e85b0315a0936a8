#!/bin/bash
# tcgen05.mma issue cost per instruction (scripts/probes/mma_rate_probe.cu): lane-0 style issue against the elect.sync pattern
P=scripts/probes/mma_rate_probe
for args in "128 16 512 1 3" "128 16 512 1 4" "128 64 512 1 4" "128 128 512 1 4" "128 256 512 1 4" "64 16 512 1 4" "128 24 512 1 4"; do
  $P $args
done
