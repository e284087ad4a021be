#!/bin/bash
# round-2 call 44: two CTC epilogue groups -- slow the MMA thread / the convert warps: which side exposes the differences
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_MMA_DELAY=2000" "OAR_DBG_CONV_DELAY=2000" "OAR_DBG_CTC_GROUPS1=1 OAR_DBG_CONV_DELAY=2000" "OAR_DBG_CTC_GROUPS1=1 OAR_DBG_CONV_DELAY=6000" "X=1"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/stress_determinism.py sleep 160 2>&1 | grep -E "baseline|mismatches"
done > gpurun_out/r2c44_stress.txt 2>&1
cat gpurun_out/r2c44_stress.txt
