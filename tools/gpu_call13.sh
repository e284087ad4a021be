#!/bin/bash
# round-2 call 14: two-lane recognition -- A/B bench, then the pipeline-level tests
set -x
mkdir -p gpurun_out
OAR_DBG_ONE_STREAM=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c14_bench_one.json 2> gpurun_out/r2c14_bench_one.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c14_bench_two.json 2> gpurun_out/r2c14_bench_two.err
python - <<'P'
import json
for f in ("one","two"):
    d=json.loads(open(f"gpurun_out/r2c14_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), d["stage_ms_last_step"])
P
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py tests/test_line_orientation.py -m gpu -q > gpurun_out/r2c14_tests.log 2>&1
tail -6 gpurun_out/r2c14_tests.log
