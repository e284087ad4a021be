#!/bin/bash
# round-2 call 60: evidence for the round's final state, part 1 -- suite, per-layer profile, benches with the CPU port
# beside them, the reference arm, the ncu launch list of the bench command
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2f_smi.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2f_gpu_tests.log 2>&1
tail -3 gpurun_out/r2f_gpu_tests.log
timeout 300 python tools/layerprof.py --out gpurun_out/r2f_layerprof.json > gpurun_out/r2f_layerprof.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
cp gpurun_out/bench_kernels.json gpurun_out/r2f_bench_kernels.json
timeout 900 python bench.py --workload rec512 > gpurun_out/r2f_bench_rec512.json 2> gpurun_out/r2f_bench_rec512.err
timeout 900 python bench.py --workload layout --steps 10 > gpurun_out/r2f_bench_layout.json 2> gpurun_out/r2f_bench_layout.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_launch_bench.log 2>&1
ls -la gpurun_out | grep r2f
