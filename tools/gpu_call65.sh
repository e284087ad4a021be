#!/bin/bash
# round-2 call 65: evidence part 2 -- one full ncu capture per kernel function at the bench configuration
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-id ::regex:.:2 -o gpurun_out/r2f_classes -f \
    python tools/ncu_step.py > gpurun_out/r2f_ncu_classes.log 2>&1
tail -2 gpurun_out/r2f_ncu_classes.log; ls -la gpurun_out/r2f_classes.ncu-rep
