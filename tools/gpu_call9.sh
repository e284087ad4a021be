#!/bin/bash
# round-2 call 9: HGNetV2 / server recogniser / layout encoder on the CUDA engine
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_server_models.py tests/test_zz_baseline_configs.py -m gpu -q -x > gpurun_out/r2c9_server_tests.log 2>&1
tail -30 gpurun_out/r2c9_server_tests.log
