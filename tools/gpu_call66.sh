#!/bin/bash
# round-2 call 66: evidence part 3 -- the 6th launch of every tensor-core kernel function (recogniser shapes) and the
# CTC head / fold conv
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --kernel-id ::regex:'lcblock|conv_rowtaps|conv_gemm|conv_halo|dwconv':6 -o gpurun_out/r2f_classes6 -f \
    python tools/ncu_step.py > gpurun_out/r2f_ncu_classes6.log 2>&1
tail -2 gpurun_out/r2f_ncu_classes6.log; ls -la gpurun_out/r2f_classes6.ncu-rep
