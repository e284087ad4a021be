#!/bin/bash
# round-2 call 34: which lcblock_tc variant is the nondeterministic one under two lanes
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_FB_OFF=1" "OAR_DBG_FB_OFF=2" "OAR_DBG_FB_OFF=4" "OAR_DBG_FB_OFF=8" "OAR_DBG_FB_OFF=12" "OAR_DBG_FB_OFF=3" "OAR_DBG_FB_NS=2" "X=1"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/det_diff.py 6 2>&1 | grep -E "^run|regions"
done > gpurun_out/r2c34_diff.txt 2>&1
cat gpurun_out/r2c34_diff.txt
