#!/bin/bash
# round-2 call 42: two CTC epilogue groups -- does a pause between the accumulator-full wait and the first TMEM read cure it?
set -x
mkdir -p gpurun_out
for mode in "X=1" "OAR_DBG_CTC_DELAY=300" "OAR_DBG_CTC_DELAY=3000" "OAR_DBG_CTC_GROUPS1=1" "OAR_DBG_CTC_STREAM=1"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/stress_determinism.py sleep 120 2>&1 | grep -E "baseline|mismatches"
done > gpurun_out/r2c42_stress.txt 2>&1
cat gpurun_out/r2c42_stress.txt
