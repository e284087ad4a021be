#!/bin/bash
# round-2 call 28: resident CTC-head weights; third weight stage for the shared-A blocks
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c28_tests.log 2>&1
tail -4 gpurun_out/r2c28_tests.log
for rep in 1 2; do
  timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "full_batch or jitter" > gpurun_out/r2c28_t$rep.log 2>&1
  echo "rep $rep: $(tail -n 1 gpurun_out/r2c28_t$rep.log)"
done
OAR_DBG_TILES=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c28_layerprof.json > gpurun_out/r2c28_layerprof.txt 2>&1
grep -E "share 1" gpurun_out/r2c28_layerprof.txt | sort | uniq -c
grep -E "^(lcblock5|ctc_head|total)" gpurun_out/r2c28_layerprof.txt
OAR_DBG_CTC_STREAM=1 OAR_FB_SHARE_NB=2 timeout 300 python tools/layerprof.py --out gpurun_out/r2c28_layerprof_old.json > gpurun_out/r2c28_layerprof_old.txt 2>&1
grep -E "^(lcblock5|ctc_head|total)" gpurun_out/r2c28_layerprof_old.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c28_bench.json 2> gpurun_out/r2c28_bench.err
timeout 600 python bench.py --workload rec512 --no-cpu-baseline > gpurun_out/r2c28_bench_rec512.json 2> gpurun_out/r2c28_bench_rec512.err
python - <<'P'
import json
for f in ("bench","bench_rec512"):
    d=json.loads(open(f"gpurun_out/r2c28_{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
