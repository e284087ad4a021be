#!/bin/bash
# round-2 call 58: folded 3x3 conv with a static MMA issue loop
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py tests/test_zz_baseline_configs.py -m gpu -q -x > gpurun_out/r2c58_tests.log 2>&1
tail -3 gpurun_out/r2c58_tests.log
for ks in 2 1; do
OAR_DBG_FOLD_KS=$ks timeout 300 python tools/layerprof.py --out gpurun_out/r2c58_layerprof_ks$ks.json > gpurun_out/r2c58_layerprof_ks$ks.txt 2>&1
grep -E "^(convkxk|total)" gpurun_out/r2c58_layerprof_ks$ks.txt | head -4
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c58_bench.json 2> gpurun_out/r2c58_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c58_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
