#!/bin/bash
# round-2 call 4: stem_u8 v2 / deconv_pair v2 / se_fc x4 images -- GPU tests, per-layer profile, ncu of the new kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c4_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c4_gpu_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c4_layerprof.json > gpurun_out/r2c4_layerprof.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stem_u8|deconv_pair|se_fc|se_gap' -c 10 -o gpurun_out/r2c4_simt -f \
    python tools/ncu_step.py > gpurun_out/r2c4_ncu.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
tail -c 1200 gpurun_out/r2c4_bench.json
