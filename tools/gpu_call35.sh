#!/bin/bash
# round-2 call 35: two-lane nondeterminism of the plain 1x1 kernel -- one convert team / one staging tile / lanes serialised
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_FB_ONE_TEAM=1" "OAR_DBG_FB_EP=1" "OAR_DBG_LANE_SERIAL=1" "OAR_DBG_FB_OFF=8" "OAR_DBG_FB_OFF=8 OAR_DBG_FB_ONE_TEAM=1" "OAR_DBG_FB_OFF=8 OAR_DBG_FB_EP=1"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/det_diff.py 6 2>&1 | grep -E "^run|regions"
done > gpurun_out/r2c35_diff.txt 2>&1
cat gpurun_out/r2c35_diff.txt
