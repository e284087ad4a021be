#!/bin/bash
# round-2 call 10: layout detector end to end on the device
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_server_models.py -m gpu -q -x -k "layout" > gpurun_out/r2c10_layout_tests.log 2>&1
tail -40 gpurun_out/r2c10_layout_tests.log
