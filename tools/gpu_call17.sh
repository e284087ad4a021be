#!/bin/bash
# round-2 call 17: conv_halo_tc -- correctness first (nets vs oracle), then benches
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_server_models.py tests/test_zz_baseline_configs.py -m gpu -q -x > gpurun_out/r2c17_tests.log 2>&1
tail -12 gpurun_out/r2c17_tests.log
OAR_DBG_TILES=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c17_layerprof.json > gpurun_out/r2c17_layerprof.txt 2>&1
grep -E "convkxk|total" gpurun_out/r2c17_layerprof.txt
grep halo gpurun_out/r2c17_layerprof.txt | sort | uniq -c
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
timeout 600 python bench.py --workload layout --no-cpu-baseline --steps 10 > gpurun_out/r2c17_bench_layout.json 2> gpurun_out/r2c17_bench_layout.err
python - <<'P'
import json
for f in ("bench","bench_layout"):
    d=json.loads(open(f"gpurun_out/r2c17_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3))
    for k in d["top_kernels"][:6]: print("   ", k["name"], k["launches_per_step"], round(k["ms_per_step"],3), k["bound"], round(k["roofline_frac"],3))
P
