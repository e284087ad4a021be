#!/bin/bash
# round-2 call 38: two-lane nondeterminism -- the persistent kernel asks for all of the SM's shared memory (no co-resident CTA)
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_FB_SMEMMAX=1" "OAR_DBG_FB_OFF=8 OAR_DBG_FB_SMEMMAX=1" "OAR_DBG_FB_OFF=8"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/det_diff.py 8 2>&1 | grep -E "^run|regions"
done > gpurun_out/r2c38_diff.txt 2>&1
cat gpurun_out/r2c38_diff.txt
