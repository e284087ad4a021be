#!/bin/bash
# round-2 call 70: labelling passes at four pixels per thread, no background label writes -- suite, profile, bench
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c70_tests.log 2>&1
tail -3 gpurun_out/r2c70_tests.log
OAR_DBG_POISON=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "db_post or pipeline or det" > gpurun_out/r2c70_tests_poison.log 2>&1
tail -2 gpurun_out/r2c70_tests_poison.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c70_lp.json > gpurun_out/r2c70_lp.txt 2>&1
grep -E "^(db_|total)" gpurun_out/r2c70_lp.txt | cut -c1-110
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c70_bench.json 2> gpurun_out/r2c70_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c70_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
