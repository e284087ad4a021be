#!/bin/bash
# round-2 call 33: which kernel family makes the two-lane recognition nondeterministic
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_ENGINE_CAP=1" "OAR_DBG_NOHALO=1" "OAR_DBG_NOCTC=1" "OAR_DBG_NOHALO=1 OAR_DBG_NOCTC=1" "OAR_DBG_NOSIMTFUSE=1" "OAR_DBG_RESIZE2=1"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/det_diff.py 6 2>&1 | grep -E "^run|regions"
done > gpurun_out/r2c33_diff.txt 2>&1
cat gpurun_out/r2c33_diff.txt
