#!/bin/bash
# round-2 call 8: staged pipeline refactor + oar_crop_rec_run / oar_rec_run_ex / oar_pipeline_run_multi, bench modes
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c8_gpu_tests.log 2>&1
tail -15 gpurun_out/r2c8_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
tail -c 2500 gpurun_out/r2c8_bench.json
timeout 600 python bench.py --workload rec512 > gpurun_out/r2c8_bench_rec512.json 2> gpurun_out/r2c8_bench_rec512.err
tail -c 2500 gpurun_out/r2c8_bench_rec512.json
tail -5 gpurun_out/r2c8_bench_rec512.err
