#!/bin/bash
# round-2 call 37: two-lane nondeterminism of the dwconv -> plain 1x1 pairs -- a fence in the producer / a spacer kernel
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_FB_FENCE=1" "OAR_DBG_FB_PAD=1" "OAR_DBG_FB_PAD=20000" "OAR_DBG_FB_OFF=8 OAR_DBG_FB_FENCE=1" "OAR_DBG_FB_OFF=8 OAR_DBG_FB_PAD=20000" "OAR_DBG_FB_OFF=8"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/det_diff.py 6 2>&1 | grep -E "^run|regions"
done > gpurun_out/r2c37_diff.txt 2>&1
cat gpurun_out/r2c37_diff.txt
