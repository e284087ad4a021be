#!/bin/bash
# round-2 call 2: stem_u8 / deconv_pair / se_fc / ctc_combine changes -- GPU tests, bench, per-layer profile
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c2_gpu_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c2_layerprof.json > gpurun_out/r2c2_layerprof.txt 2>&1
timeout 600 python bench.py > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
tail -c 1500 gpurun_out/r2c2_bench.json
