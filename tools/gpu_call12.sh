#!/bin/bash
# round-2 call 12: the whole GPU suite on the current tree
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c12_gpu_tests.log 2>&1
tail -8 gpurun_out/r2c12_gpu_tests.log
