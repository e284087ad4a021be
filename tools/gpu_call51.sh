#!/bin/bash
# round-2 call 51: conv_halo_tc with resident weights, enabled for the detector's narrow 96 -> 24 convs
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c51_tests.log 2>&1
tail -3 gpurun_out/r2c51_tests.log
OAR_DBG_TILES=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c51_layerprof.json > gpurun_out/r2c51_layerprof.txt 2>&1
grep "halo" gpurun_out/r2c51_layerprof.txt | sort | uniq -c
grep -E "^(convkxk|total)" gpurun_out/r2c51_layerprof.txt
OAR_DBG_HALO_STREAM=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c51_layerprof_old.json > gpurun_out/r2c51_layerprof_old.txt 2>&1
grep -E "^(convkxk|total)" gpurun_out/r2c51_layerprof_old.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c51_bench.json 2> gpurun_out/r2c51_bench.err
timeout 600 python bench.py --workload layout --no-cpu-baseline --steps 10 > gpurun_out/r2c51_bench_layout.json 2> gpurun_out/r2c51_bench_layout.err
python - <<'P'
import json
for f in ("bench","bench_layout"):
    d=json.loads(open(f"gpurun_out/r2c51_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
