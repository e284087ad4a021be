#!/bin/bash
# round-2 call 24: which change made the 32-page run nondeterministic?
set -x
mkdir -p gpurun_out
for mode in "OAR_DBG_NOHALO=1" "OAR_DBG_ONE_STREAM=1" "OAR_DBG_NOHALO=1 OAR_DBG_ONE_STREAM=1" "X=1"; do
  for rep in 1 2; do
    env $mode timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "full_batch or jitter" > gpurun_out/r2c24_t.log 2>&1
    echo "mode [$mode] rep $rep: $(tail -1 gpurun_out/r2c24_t.log)"
  done
done
