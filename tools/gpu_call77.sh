#!/bin/bash
# round-2 call 77: soak of the determinism fix on the final build (two lanes, two CTC groups: the defaults), then the
# final bench lines
set -x
mkdir -p gpurun_out
timeout 600 python tools/det_diff.py 40 2>&1 | grep -E "^run|regions" | awk '{print $3,$4,$5,$6,$7}' | sort | uniq -c > gpurun_out/r2f_soak_det_diff.txt; cat gpurun_out/r2f_soak_det_diff.txt
timeout 900 python tools/stress_determinism.py sleep 1000 2>&1 | grep -E "baseline|mismatches" > gpurun_out/r2f_soak_stress.txt; cat gpurun_out/r2f_soak_stress.txt
timeout 600 python tools/stress_determinism.py big 200 2>&1 | grep -E "baseline|mismatches" >> gpurun_out/r2f_soak_stress.txt; tail -1 gpurun_out/r2f_soak_stress.txt
timeout 600 python tools/ctc_dump_diff.py 500 2>&1 | tail -1 >> gpurun_out/r2f_soak_stress.txt; tail -1 gpurun_out/r2f_soak_stress.txt
timeout 300 python tools/layerprof.py --out gpurun_out/r2f_layerprof.json > gpurun_out/r2f_layerprof.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
cp gpurun_out/bench_kernels.json gpurun_out/r2f_bench_kernels.json
timeout 900 python bench.py --workload rec512 > gpurun_out/r2f_bench_rec512.json 2> gpurun_out/r2f_bench_rec512.err
timeout 900 python bench.py --workload layout --steps 10 > gpurun_out/r2f_bench_layout.json 2> gpurun_out/r2f_bench_layout.err
python - <<'P'
import json
for f in ("bench_n1","bench_rec512","bench_layout"):
    d=json.loads(open(f"gpurun_out/r2f_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), d["roofline"]["step_frac"], d["roofline"]["kernel"], round(d["roofline"]["frac"],3), round(d["roofline"]["achieved"],1), d["cpu_baseline"]["value"], d["parity_check"])
P
