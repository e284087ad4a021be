#!/bin/bash
# round-2 call 39: one recognition lane by default, two-pass CTC epilogue, resident CTC weights -- suite, determinism x3, benches
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c39_tests.log 2>&1
tail -4 gpurun_out/r2c39_tests.log
timeout 300 python tools/det_diff.py 8 2>&1 | grep -E "^run|regions" > gpurun_out/r2c39_diff.txt
cat gpurun_out/r2c39_diff.txt
timeout 300 python tools/layerprof.py --out gpurun_out/r2c39_layerprof.json > gpurun_out/r2c39_layerprof.txt 2>&1
tail -1 gpurun_out/r2c39_layerprof.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c39_bench.json 2> gpurun_out/r2c39_bench.err
OAR_REC_LANES=2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c39_bench_2lanes.json 2> gpurun_out/r2c39_bench_2lanes.err
timeout 600 python bench.py --workload rec512 --no-cpu-baseline > gpurun_out/r2c39_bench_rec512.json 2> gpurun_out/r2c39_bench_rec512.err
python - <<'P'
import json
for f in ("bench","bench_2lanes","bench_rec512"):
    d=json.loads(open(f"gpurun_out/r2c39_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
