#!/bin/bash
# round-2 call 67: evidence part 4 -- every launch of the persistent kernel in one 512-crop recogniser run (3x3 / 5x5
# blocks, 1x1 convs, squeeze-excite 1x1, CTC head), the folded 3x3 conv at 240x240, and the pre/post kernels
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none -k regex:'lcblock_tc' --launch-skip 22 -c 22 -o gpurun_out/r2f_rec512_fb -f \
    python tools/ncu_step.py --rec512 --steps 2 > gpurun_out/r2f_ncu_rec512.log 2>&1
tail -2 gpurun_out/r2f_ncu_rec512.log
timeout 900 ncu --set full --clock-control none -k regex:'conv_halo' --launch-skip 3 -c 2 -o gpurun_out/r2f_fold -f \
    python tools/ncu_step.py > gpurun_out/r2f_ncu_fold.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-id ::regex:'db_|crop_|deconv_pair|normalize|cls_|rotate':1 -o gpurun_out/r2f_prepost -f \
    python tools/ncu_step.py > gpurun_out/r2f_ncu_prepost.log 2>&1
ls -la gpurun_out/r2f_rec512_fb.ncu-rep gpurun_out/r2f_fold.ncu-rep gpurun_out/r2f_prepost.ncu-rep
