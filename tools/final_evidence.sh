mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2h_smi.txt 2>&1
timeout 500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2h_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -3 gpurun_out/r2h_smoke.log
timeout 300 python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
cp gpurun_out/bench_kernels.json gpurun_out/r2h_bench_kernels.json
timeout 300 python bench.py --workload rec512 > gpurun_out/r2h_bench_rec512.json 2> gpurun_out/r2h_bench_rec512.err
timeout 300 python bench.py --workload layout --steps 10 > gpurun_out/r2h_bench_layout.json 2> gpurun_out/r2h_bench_layout.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
timeout 300 python tools/layerprof.py --out gpurun_out/r2h_layerprof.json > gpurun_out/r2h_layerprof.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_ncu_bench.log 2>&1
python - <<PY
import json
for f in ("r2h_bench_n1","r2h_bench_rec512","r2h_bench_layout","r2h_bench_reference"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), d.get("gpu_launches"), d.get("host_submissions_per_step"), (d.get("roofline") or {}).get("step_frac"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), d.get("parity_check"))
    except Exception as e: print(f, "ERR", e)
PY
wc -l gpurun_out/r2h_launches.csv; tail -2 gpurun_out/r2h_bench_rec512.err gpurun_out/r2h_bench_layout.err
