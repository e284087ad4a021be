#!/bin/bash
# round-2 call 71 (2 GPUs): smoke(), the GPU suite twice more, the 2-GPU bench under torchrun, the in-process 2-GPU test
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c71_smoke.log 2>&1; tail -2 gpurun_out/r2c71_smoke.log
for rep in 1 2; do
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c71_tests$rep.log 2>&1; tail -1 gpurun_out/r2c71_tests$rep.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c71_bench_n2.json 2> gpurun_out/r2c71_bench_n2.err
tail -c 900 gpurun_out/r2c71_bench_n2.json | head -c 900; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2c71_bench_ref_n2.json 2> gpurun_out/r2c71_bench_ref_n2.err
tail -c 300 gpurun_out/r2c71_bench_ref_n2.json; echo
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c71_bench_n2.json").read().strip().splitlines()[-1])
print("n2:", d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), d["scaling"])
P
