#!/bin/bash
# round-2 call 68: evidence part 4 (summaries made on the box; the reports stay there) -- every launch of the persistent
# kernel in one 512-crop recogniser run, the folded 3x3 conv at 240x240, the pre/post kernels
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none -k regex:'lcblock_tc' --launch-skip 22 -c 22 -o /tmp/r2f_rec512_fb -f \
    python tools/ncu_step.py --rec512 --steps 2 > gpurun_out/r2f_ncu_rec512.log 2>&1
(echo "# every lcblock_tc launch of one oar_rec_run over 512 synthetic 48x320 crops (configs[2]), second run: 3x3 blocks, 5x5 blocks, 1x1 convs (squeeze-excite, 480-channel, SVTR neck), CTC head (last row, grid 144)"; python tools/ncu_table.py /tmp/r2f_rec512_fb.ncu-rep) > gpurun_out/r2f_rec512_lcblock.txt
timeout 900 ncu --set full --clock-control none -k regex:'conv_halo' --launch-skip 3 -c 2 -o /tmp/r2f_fold -f \
    python tools/ncu_step.py > gpurun_out/r2f_ncu_fold.log 2>&1
(echo "# conv_halo_tc in fold mode: the detector's 3x3 96 -> 24 convs at 240x240 (neck level 0, DBHead), 32 pages"; python tools/ncu_table.py /tmp/r2f_fold.ncu-rep) > gpurun_out/r2f_fold_conv.txt
timeout 900 ncu --set full --clock-control none --kernel-id ::regex:'db_|crop_|deconv_pair|stem_u8':1 -o /tmp/r2f_prepost -f \
    python tools/ncu_step.py > gpurun_out/r2f_ncu_prepost.log 2>&1
(echo "# first launch of the crop / DB post-process / deconv / stem kernels in the bench step (32 pages 960x960)"; python tools/ncu_table.py /tmp/r2f_prepost.ncu-rep) > gpurun_out/r2f_prepost.txt
cat gpurun_out/r2f_rec512_lcblock.txt gpurun_out/r2f_fold_conv.txt gpurun_out/r2f_prepost.txt | cut -c1-190
