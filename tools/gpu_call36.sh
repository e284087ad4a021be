#!/bin/bash
# round-2 call 36: which plain 1x1 launches (wide / narrow, many / few items) carry the two-lane nondeterminism
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_FB_OFF=16" "OAR_DBG_FB_OFF=32" "OAR_DBG_FB_OFF=64" "OAR_DBG_FB_OFF=128" "OAR_DBG_FB_OFF=24" "OAR_DBG_FB_OFF=72"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/det_diff.py 6 2>&1 | grep -E "^run|regions"
done > gpurun_out/r2c36_diff.txt 2>&1
cat gpurun_out/r2c36_diff.txt
