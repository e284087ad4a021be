#!/bin/bash
# round-2 call 16: ncu of conv_halo_tc at detector and layout shapes
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_tc -c 6 -o gpurun_out/r2c16_halo_det -f python tools/ncu_step.py --batch 8 > gpurun_out/r2c16_ncu_det.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_tc --launch-skip 6 -c 8 -o gpurun_out/r2c16_halo_layout -f python tools/ncu_layout_step.py > gpurun_out/r2c16_ncu_layout.log 2>&1
tail -3 gpurun_out/r2c16_ncu_det.log gpurun_out/r2c16_ncu_layout.log
