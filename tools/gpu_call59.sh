#!/bin/bash
# round-2 call 59: folded 3x3 conv, BN = 80, three input stages; full suite
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c59_tests.log 2>&1
tail -3 gpurun_out/r2c59_tests.log
OAR_DBG_TILES=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c59_layerprof.json > gpurun_out/r2c59_layerprof.txt 2>&1
grep "fold\]" gpurun_out/r2c59_layerprof.txt | sort | uniq -c | head -3
grep -E "^(convkxk|total)" gpurun_out/r2c59_layerprof.txt | head -5
timeout 300 python tools/det_diff.py 6 2>&1 | grep -E "^run|regions" | sort | uniq -c | head
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c59_bench.json 2> gpurun_out/r2c59_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c59_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
