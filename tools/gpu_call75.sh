#!/bin/bash
# round-2 call 75: dw_tma.cu with k-blocks as the fastest item index -- suite, determinism, A/B per launch
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c75_tests.log 2>&1
tail -3 gpurun_out/r2c75_tests.log
timeout 300 python tools/det_diff.py 4 2>&1 | grep -E "^run|regions" | sort | uniq -c | head -5
for mode in "X=1" "OAR_DBG_NODWTMA=1"; do
env $mode timeout 300 python tools/layerprof.py --out gpurun_out/r2c75_lp.json > gpurun_out/r2c75_lp.txt 2>&1
echo "== $mode"; grep -E "^(dwconv|se_gap)" gpurun_out/r2c75_lp.txt; tail -1 gpurun_out/r2c75_lp.txt | cut -c1-30
done
