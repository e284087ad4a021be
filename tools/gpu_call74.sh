#!/bin/bash
# round-2 call 74: stand-alone depthwise conv on TMA-staged tiles (dw_tma.cu) -- suite, determinism, A/B, bench
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c74_tests.log 2>&1
tail -6 gpurun_out/r2c74_tests.log
timeout 300 python tools/det_diff.py 5 2>&1 | grep -E "^run|regions" | sort | uniq -c | head -6
for mode in "X=1" "OAR_DBG_NODWTMA=1"; do
env $mode OAR_DBG_TILES=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c74_lp.json > gpurun_out/r2c74_lp.txt 2>&1
echo "== $mode"; grep "dwtma\]" gpurun_out/r2c74_lp.txt | sort | uniq -c | head -12; grep -E "^(dwconv|se_gap)" gpurun_out/r2c74_lp.txt | awk '{a[$1]+=$5} END {for (k in a) print k, a[k]}'; tail -1 gpurun_out/r2c74_lp.txt | cut -c1-30
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c74_bench.json 2> gpurun_out/r2c74_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c74_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
