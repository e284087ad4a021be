#!/bin/bash
# round-2 call 27: lcblock_tc with one A operand for both N tiles (5x5 blocks wider than 256 channels fused again)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c27_tests.log 2>&1
tail -8 gpurun_out/r2c27_tests.log
for rep in 1 2; do
  timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "full_batch or jitter" > gpurun_out/r2c27_t$rep.log 2>&1
  echo "rep $rep: $(tail -1 gpurun_out/r2c27_t$rep.log)"
done
OAR_DBG_TILES=1 timeout 300 python tools/layerprof.py --out gpurun_out/r2c27_layerprof.json > gpurun_out/r2c27_layerprof.txt 2>&1
grep -E "share 1" gpurun_out/r2c27_layerprof.txt | sort | uniq -c
grep -E "^(lcblock|dwconv|pwconv|se_pwconv|total)" gpurun_out/r2c27_layerprof.txt
OAR_FB_SHARE=2 timeout 300 python tools/layerprof.py --out gpurun_out/r2c27_layerprof_s2.json > gpurun_out/r2c27_layerprof_s2.txt 2>&1
grep -E "^(pwconv|se_pwconv|total)" gpurun_out/r2c27_layerprof_s2.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c27_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'resize_fused|attn_core_split|conv_halo' -c 6 -o gpurun_out/r2c27_small -f \
    python tools/ncu_step.py > gpurun_out/r2c27_ncu.log 2>&1
tail -3 gpurun_out/r2c27_ncu.log
