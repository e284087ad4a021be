#!/bin/bash
# round-2 call 72: conv_halo_tc general loop with one issue block per weight stage -- suite, layout / pipeline benches
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c72_tests.log 2>&1
tail -3 gpurun_out/r2c72_tests.log
timeout 600 python bench.py --workload layout --no-cpu-baseline --steps 10 > gpurun_out/r2c72_bench_layout.json 2> gpurun_out/r2c72_bench_layout.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c72_bench.json 2> gpurun_out/r2c72_bench.err
python - <<'P'
import json
for f in ("bench_layout","bench"):
    d=json.loads(open(f"gpurun_out/r2c72_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
    for k in d["top_kernels"][:4]: print("   ", k["name"], k["launches_per_step"], round(k["ms_per_step"],3), k["bound"], round(k["roofline_frac"],3))
P
