#!/bin/bash
# round-2 call 64: residual squeeze-excite applied by its upadd / upsample consumer (RSEFPN) -- suite, A/B
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c64_tests.log 2>&1
tail -3 gpurun_out/r2c64_tests.log
for mode in "X=1" "OAR_DBG_NOLAZYSE=1"; do
env $mode timeout 300 python tools/layerprof.py --out gpurun_out/r2c64_lp.json > gpurun_out/r2c64_lp.txt 2>&1
echo "== $mode"; grep -E "^(se_apply|upadd|upsample_into)" gpurun_out/r2c64_lp.txt | awk '{a[$1]+=$5} END {for (k in a) print k, a[k]}'; tail -1 gpurun_out/r2c64_lp.txt | cut -c1-120
done
