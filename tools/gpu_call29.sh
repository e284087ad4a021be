#!/bin/bash
# round-2 call 29: padded attention rows; full ncu captures (with source) of the persistent kernel's launches in one
# 512-crop recogniser run: CTC head, 5x5 blocks, shared-A blocks
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rec or pipeline" > gpurun_out/r2c29_tests.log 2>&1
tail -3 gpurun_out/r2c29_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c29_layerprof.json > gpurun_out/r2c29_layerprof.txt 2>&1
grep -E "^(attn_core|total)" gpurun_out/r2c29_layerprof.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'lcblock_tc' --launch-skip 34 -c 34 -o gpurun_out/r2c29_fb -f \
    python tools/ncu_step.py --rec512 > gpurun_out/r2c29_ncu.log 2>&1
tail -3 gpurun_out/r2c29_ncu.log
