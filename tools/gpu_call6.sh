#!/bin/bash
# round-2 call 6: dwconv_tma 2 teams x 2 stages, stem_u8 v4 -- GPU tests, per-layer profile, ncu at recogniser sizes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c6_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c6_gpu_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c6_layerprof.json > gpurun_out/r2c6_layerprof.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dwconv_tma|stem_u8|deconv_pair' -c 8 -o gpurun_out/r2c6_dw -f \
    python tools/recprof.py > gpurun_out/r2c6_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_u8|deconv_pair' -c 2 -o gpurun_out/r2c6_det -f \
    python tools/ncu_step.py > gpurun_out/r2c6_ncu2.log 2>&1
