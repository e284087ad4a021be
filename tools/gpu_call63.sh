#!/bin/bash
# round-2 call 63: third A operand buffer in lcblock_tc where it fits -- suite, determinism, A/B
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c63_tests.log 2>&1
tail -3 gpurun_out/r2c63_tests.log
timeout 300 python tools/det_diff.py 6 2>&1 | grep -E "^run|regions" | sort | uniq -c | head -8
timeout 300 python tools/stress_determinism.py sleep 150 2>&1 | grep -E "baseline|mismatches"
for na in 3 2; do
OAR_DBG_TILES=1 OAR_FB_NA=$na timeout 300 python tools/layerprof.py --out gpurun_out/r2c63_lp_$na.json > gpurun_out/r2c63_lp_$na.txt 2>&1
echo "== na $na"; grep -E "^(lcblock3|lcblock5|pwconv_tc|se_pwconv|ctc_head)" gpurun_out/r2c63_lp_$na.txt | awk '{a[$1]+=$5} END {for (k in a) print k, a[k]}'; tail -1 gpurun_out/r2c63_lp_$na.txt | cut -c1-30
done
grep "fused\]" gpurun_out/r2c63_lp_3.txt | grep -c "abufs 3"; grep "fused\]" gpurun_out/r2c63_lp_3.txt | grep -c "abufs 2"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c63_bench.json 2> gpurun_out/r2c63_bench.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2c63_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), "step_frac", round(d["roofline"]["step_frac"],3))
P
