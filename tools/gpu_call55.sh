#!/bin/bash
# round-2 call 55: narrow 3x3 convs with the kernel columns folded into N; dwconv at 4 CTAs per SM
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c55_tests.log 2>&1
tail -12 gpurun_out/r2c55_tests.log
OAR_DBG_TILES=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c55_layerprof.json > gpurun_out/r2c55_layerprof.txt 2>&1
grep "fold\]" gpurun_out/r2c55_layerprof.txt | sort | uniq -c
grep -E "^(convkxk|dwconv|total)" gpurun_out/r2c55_layerprof.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c55_bench.json 2> gpurun_out/r2c55_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c55_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
