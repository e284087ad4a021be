#!/bin/bash
# round-2 call 41: second CTC epilogue group -- parity, determinism (8 runs + jitter test x3), timing
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c41_tests.log 2>&1
tail -3 gpurun_out/r2c41_tests.log
timeout 300 python tools/det_diff.py 10 2>&1 | grep -E "^run|regions" > gpurun_out/r2c41_diff.txt
cat gpurun_out/r2c41_diff.txt
for rep in 1 2 3; do
  timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "full_batch or jitter" > gpurun_out/r2c41_t$rep.log 2>&1
  echo "rep $rep: $(tail -n 1 gpurun_out/r2c41_t$rep.log)"
done
timeout 300 python tools/stress_determinism.py > gpurun_out/r2c41_stress.txt 2>&1; tail -3 gpurun_out/r2c41_stress.txt
timeout 300 python tools/layerprof.py --out gpurun_out/r2c41_layerprof.json > gpurun_out/r2c41_layerprof.txt 2>&1
grep -E "^(ctc_head|total)" gpurun_out/r2c41_layerprof.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c41_bench.json 2> gpurun_out/r2c41_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c41_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
