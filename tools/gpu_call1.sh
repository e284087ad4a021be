#!/bin/bash
# round-2 call 1: baseline evidence. One full ncu capture per kernel function at the bench configuration, the launch
# list, the per-layer profile, a racecheck/synccheck pass over the recogniser head (one- and two-group CTC epilogue).
set -x
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt
timeout 300 python tools/layerprof.py --out gpurun_out/r2c1_layerprof.json > gpurun_out/r2c1_layerprof.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-id ::regex:.:2 -o gpurun_out/r2c1_classes -f \
    python tools/ncu_step.py > gpurun_out/r2c1_ncu_classes.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-id ::regex:'lcblock|conv_rowtaps|conv_gemm|dwconv':6 -o gpurun_out/r2c1_classes6 -f \
    python tools/ncu_step.py > gpurun_out/r2c1_ncu_classes6.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2c1_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c1_launch_bench.log 2>&1
OAR_CTC_GROUPS=2 timeout 300 compute-sanitizer --tool synccheck python tools/recprof.py --n 64 > gpurun_out/r2c1_synccheck_g2.log 2>&1
OAR_CTC_GROUPS=2 timeout 400 compute-sanitizer --tool racecheck python tools/recprof.py --n 64 > gpurun_out/r2c1_racecheck_g2.log 2>&1
OAR_CTC_GROUPS=2 timeout 200 python tools/stress_determinism.py sleep 300 > gpurun_out/r2c1_stress_g2.log 2>&1
ls -la gpurun_out
