#!/bin/bash
# round-2 call 5: dwconv_tma + se_gap v2 + se_fc<IMG> -- GPU tests, per-layer profile, ncu of the new kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2c5_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c5_gpu_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c5_layerprof.json > gpurun_out/r2c5_layerprof.txt 2>&1
OAR_DBG_TILES=1 timeout 300 python tools/recprof.py > gpurun_out/r2c5_tiles.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dwconv_tma|se_fc|se_gap|stem_u8' -c 14 -o gpurun_out/r2c5_dw -f \
    python tools/ncu_step.py > gpurun_out/r2c5_ncu.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
tail -c 1200 gpurun_out/r2c5_bench.json
