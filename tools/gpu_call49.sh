#!/bin/bash
# round-2 call 49: detector network in two halves while page uploads are in flight (e2e) -- suite, bench A/B
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c49_tests.log 2>&1
tail -3 gpurun_out/r2c49_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c49_bench.json 2> gpurun_out/r2c49_bench.err
OAR_DET_SPLIT_MIN=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c49_bench_nosplit.json 2> gpurun_out/r2c49_bench_nosplit.err
python - <<'P'
import json
for f in ("bench","bench_nosplit"):
    d=json.loads(open(f"gpurun_out/r2c49_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d.get("e2e_stage_ms"))
P
