#!/bin/bash
# round-2 call 56: full ncu capture (with source) of the folded 3x3 conv at 240x240
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_halo' --launch-skip 3 -c 1 -o gpurun_out/r2c56_fold -f \
    python tools/ncu_step.py > gpurun_out/r2c56_ncu.log 2>&1
tail -2 gpurun_out/r2c56_ncu.log
