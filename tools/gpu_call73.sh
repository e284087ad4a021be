#!/bin/bash
# round-2 call 73 (8 GPUs): scaling of the final build, 4 and 8 ranks under torchrun
set -x
mkdir -p gpurun_out
for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c73_bench_n$n.json 2> gpurun_out/r2c73_bench_n$n.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2c73_bench_n$n.json").read().strip().splitlines()[-1])
print("n$n:", d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), d["scaling"], d["clocks"])
P
done
