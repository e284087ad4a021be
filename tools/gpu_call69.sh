#!/bin/bash
# round-2 call 69: deconv_pair without bank conflicts, warp-per-image compaction -- suite, per-layer profile, bench
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c69_tests.log 2>&1
tail -3 gpurun_out/r2c69_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c69_lp.json > gpurun_out/r2c69_lp.txt 2>&1
grep -E "^(deconv_pair|db_compact|total)" gpurun_out/r2c69_lp.txt | cut -c1-120
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c69_bench.json 2> gpurun_out/r2c69_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c69_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
