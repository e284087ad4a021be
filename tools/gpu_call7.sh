#!/bin/bash
# round-2 call 7: state to commit (stem_u8 v3, deconv_pair v2, se_gap v2, se_fc<IMG>) + the round-2 parity tests
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c7_gpu_tests.log 2>&1
tail -15 gpurun_out/r2c7_gpu_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c7_layerprof.json > gpurun_out/r2c7_layerprof.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
tail -c 1500 gpurun_out/r2c7_bench.json
