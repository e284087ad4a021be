#!/bin/bash
# round-2 call 30: full ncu captures (with source) of the persistent kernel's launches in the second of two 512-crop
# recogniser runs: CTC head, 3x3 / 5x5 blocks, 1x1 convs
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'lcblock_tc' --launch-skip 22 -c 22 -o gpurun_out/r2c30_fb -f \
    python tools/ncu_step.py --rec512 --steps 2 > gpurun_out/r2c30_ncu.log 2>&1
tail -3 gpurun_out/r2c30_ncu.log
