#!/bin/bash
# round-2 call 26: fused resize, register layernorm, split-key attention -- parity suite, per-layer profile, bench
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c26_tests.log 2>&1
tail -8 gpurun_out/r2c26_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c26_layerprof.json > gpurun_out/r2c26_layerprof.txt 2>&1
grep -E "resize|layernorm|attn_core|total" gpurun_out/r2c26_layerprof.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c26_bench.json 2> gpurun_out/r2c26_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c26_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
