#!/bin/bash
# round-2 call 32: what differs between two runs of the 32-page step, under which switches
set -x
mkdir -p gpurun_out
for mode in "X=1" "OAR_DBG_CTC_STREAM=1" "OAR_DBG_CTC_EPI1=1" "OAR_DBG_ONE_STREAM=1" "OAR_DBG_CTC_EPI1=1 OAR_DBG_CTC_STREAM=1"; do
  echo "== mode [$mode]"
  env $mode timeout 300 python tools/det_diff.py 4 2>&1 | tail -18
done > gpurun_out/r2c32_diff.txt 2>&1
cat gpurun_out/r2c32_diff.txt
