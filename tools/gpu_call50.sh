#!/bin/bash
# round-2 call 50: how many parts for the detector split under upload (e2e)
set -x
mkdir -p gpurun_out
for parts in 2 3 4; do
OAR_DET_SPLIT_PARTS=$parts timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c50_bench_p$parts.json 2> gpurun_out/r2c50_bench_p$parts.err
done
python - <<'P'
import json
for f in ("p2","p3","p4"):
    d=json.loads(open(f"gpurun_out/r2c50_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3))
P
