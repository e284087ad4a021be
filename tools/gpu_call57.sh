#!/bin/bash
# round-2 call 57: folded 3x3 conv with two partial accumulators (independent MMA chains)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py tests/test_zz_baseline_configs.py -m gpu -q -x > gpurun_out/r2c57_tests.log 2>&1
tail -3 gpurun_out/r2c57_tests.log
OAR_DBG_TILES=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c57_layerprof.json > gpurun_out/r2c57_layerprof.txt 2>&1
grep "fold\]" gpurun_out/r2c57_layerprof.txt | sort | uniq -c
grep -E "^(convkxk|total)" gpurun_out/r2c57_layerprof.txt | head -5
OAR_DBG_FOLD_KS=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c57_layerprof_ks1.json > gpurun_out/r2c57_layerprof_ks1.txt 2>&1
grep -E "^(convkxk|total)" gpurun_out/r2c57_layerprof_ks1.txt | head -5
