#!/bin/bash
for cfg in "32 96 0" "32 96 3" "32 96 6" "32 192 0" "32 64 4"; do
  set -- $cfg
  echo "== stage_kb=$1 ring_kb=$2 l2_ahead=$3"
  ZB_GEMV_STAGE_KB=$1 ZB_GEMV_RING_KB=$2 ZB_GEMV_L2_AHEAD=$3 python scripts/kernel_times.py 2>&1 | grep "rows=2"
done
