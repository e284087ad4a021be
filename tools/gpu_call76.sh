#!/bin/bash
# round-2 call 76: dw_tma.cu on the strided layers only -- suite, determinism, bench
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c76_tests.log 2>&1
tail -3 gpurun_out/r2c76_tests.log
OAR_DWTMA=2 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_line_orientation.py -m gpu -q -x > gpurun_out/r2c76_tests_all.log 2>&1
tail -2 gpurun_out/r2c76_tests_all.log
timeout 300 python tools/det_diff.py 4 2>&1 | grep -E "^run|regions" | sort | uniq -c | head -5
timeout 300 python tools/layerprof.py --out gpurun_out/r2c76_lp.json > gpurun_out/r2c76_lp.txt 2>&1
grep -E "^(dwconv|se_gap)" gpurun_out/r2c76_lp.txt | awk '{a[$1]+=$5} END {for (k in a) print k, a[k]}'; tail -1 gpurun_out/r2c76_lp.txt | cut -c1-30
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c76_bench.json 2> gpurun_out/r2c76_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c76_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
