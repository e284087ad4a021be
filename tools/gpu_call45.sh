#!/bin/bash
# round-2 call 45: default build after the race hunts (one lane, one CTC group, two-pass epilogue) -- suite, determinism, stress, benches
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c45_tests.log 2>&1
tail -3 gpurun_out/r2c45_tests.log
timeout 300 python tools/det_diff.py 8 2>&1 | grep -E "^run|regions" > gpurun_out/r2c45_diff.txt; cat gpurun_out/r2c45_diff.txt
timeout 300 python tools/stress_determinism.py sleep 200 2>&1 | grep -E "baseline|mismatches"
timeout 600 python bench.py > gpurun_out/r2c45_bench.json 2> gpurun_out/r2c45_bench.err
timeout 600 python bench.py --workload rec512 > gpurun_out/r2c45_bench_rec512.json 2> gpurun_out/r2c45_bench_rec512.err
timeout 600 python bench.py --workload layout --steps 10 > gpurun_out/r2c45_bench_layout.json 2> gpurun_out/r2c45_bench_layout.err
python - <<'P'
import json
for f in ("bench","bench_rec512","bench_layout"):
    d=json.loads(open(f"gpurun_out/r2c45_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3), "cpu", d.get("cpu_baseline",{}) and d["cpu_baseline"].get("value"), "parity", d.get("parity_check"))
P
