#!/bin/bash
# round-2 call 11: nvJPEG ingest tests, layout bench workload, full GPU suite
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jpeg_ingest.py -m gpu -q > gpurun_out/r2c11_jpeg_tests.log 2>&1
tail -25 gpurun_out/r2c11_jpeg_tests.log
timeout 900 python bench.py --workload layout --steps 5 --warmup 3 > gpurun_out/r2c11_bench_layout.json 2> gpurun_out/r2c11_bench_layout.err
tail -c 3000 gpurun_out/r2c11_bench_layout.json; tail -5 gpurun_out/r2c11_bench_layout.err
