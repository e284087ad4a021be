#!/bin/bash
# round-2 call 23: lcblock_tc with whole-warp issue -- parity, determinism under jitter, per-layer profile, bench
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_line_orientation.py tests/test_gpu_server_models.py -m gpu -q -x > gpurun_out/r2c23_tests.log 2>&1
tail -8 gpurun_out/r2c23_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c23_layerprof.json > gpurun_out/r2c23_layerprof.txt 2>&1
tail -1 gpurun_out/r2c23_layerprof.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c23_bench.json 2> gpurun_out/r2c23_bench.err
timeout 600 python bench.py --workload rec512 --no-cpu-baseline > gpurun_out/r2c23_bench_rec512.json 2> gpurun_out/r2c23_bench_rec512.err
timeout 600 python bench.py --workload layout --no-cpu-baseline --steps 10 > gpurun_out/r2c23_bench_layout.json 2> gpurun_out/r2c23_bench_layout.err
python - <<'P'
import json
for f in ("bench","bench_rec512","bench_layout"):
    d=json.loads(open(f"gpurun_out/r2c23_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
    for k in d["top_kernels"][:8]: print("   ", k["name"], k["launches_per_step"], round(k["ms_per_step"],3), k["bound"], round(k["roofline_frac"],3))
P
