#!/bin/bash
# round-2 call 48: two lanes + two CTC groups back on after the stage hand-back fix -- suite, determinism, stress, benches
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c48_tests.log 2>&1
tail -3 gpurun_out/r2c48_tests.log
timeout 300 python tools/det_diff.py 12 2>&1 | grep -E "^run|regions" > gpurun_out/r2c48_diff.txt; sort gpurun_out/r2c48_diff.txt | uniq -c
timeout 300 python tools/stress_determinism.py sleep 300 2>&1 | grep -E "baseline|mismatches"
timeout 300 python tools/stress_determinism.py big 100 2>&1 | grep -E "baseline|mismatches"
for rep in 1 2 3; do
  timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2c48_t$rep.log 2>&1
  echo "rep $rep: $(tail -n 1 gpurun_out/r2c48_t$rep.log)"
done
timeout 300 python tools/layerprof.py --out gpurun_out/r2c48_layerprof.json > gpurun_out/r2c48_layerprof.txt 2>&1
tail -1 gpurun_out/r2c48_layerprof.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c48_bench.json 2> gpurun_out/r2c48_bench.err
timeout 600 python bench.py --workload rec512 --no-cpu-baseline > gpurun_out/r2c48_bench_rec512.json 2> gpurun_out/r2c48_bench_rec512.err
python - <<'P'
import json
for f in ("bench","bench_rec512"):
    d=json.loads(open(f"gpurun_out/r2c48_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
