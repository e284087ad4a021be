#!/bin/bash
# round-2 call 25: HEAD re-baseline -- the whole GPU suite, the determinism tests three more times, benches
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c25_tests.log 2>&1
tail -15 gpurun_out/r2c25_tests.log
for rep in 1 2 3; do
  timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "full_batch or jitter" > gpurun_out/r2c25_t$rep.log 2>&1
  echo "rep $rep: $(tail -1 gpurun_out/r2c25_t$rep.log)"
done
timeout 300 python tools/layerprof.py --out gpurun_out/r2c25_layerprof.json > gpurun_out/r2c25_layerprof.txt 2>&1
tail -1 gpurun_out/r2c25_layerprof.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c25_bench.json 2> gpurun_out/r2c25_bench.err
tail -c 600 gpurun_out/r2c25_bench.json
