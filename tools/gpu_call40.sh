#!/bin/bash
# round-2 call 40: slimmer two-pass CTC epilogue (bias -inf for the vocabulary tail, immediate-index arg-max)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c40_tests.log 2>&1
tail -3 gpurun_out/r2c40_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2c40_layerprof.json > gpurun_out/r2c40_layerprof.txt 2>&1
grep -E "^(ctc_head|total)" gpurun_out/r2c40_layerprof.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lcblock_tc' --launch-skip 43 -c 1 -o gpurun_out/r2c40_ctc -f \
    python tools/ncu_step.py --rec512 --steps 2 > gpurun_out/r2c40_ncu.log 2>&1
tail -2 gpurun_out/r2c40_ncu.log
