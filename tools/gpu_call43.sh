#!/bin/bash
# round-2 call 43: two CTC epilogue groups -- the MMA thread confirms its own commit before publishing the accumulator
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_CTC_SELF=1" "X=1" "OAR_DBG_CTC_SELF=1 Y=2"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/stress_determinism.py sleep 160 2>&1 | grep -E "baseline|mismatches"
done > gpurun_out/r2c43_stress.txt 2>&1
cat gpurun_out/r2c43_stress.txt
