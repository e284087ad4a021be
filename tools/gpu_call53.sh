#!/bin/bash
# round-2 call 53: how sensitive are the fused blocks to the depth of their input ring (bytes in flight)?
set -x
mkdir -p gpurun_out
for ns in 0 2 3; do
  OAR_DBG_TILES=1 OAR_DBG_FB_NS=$ns timeout 300 python tools/layerprof.py --out gpurun_out/r2c53_lp_$ns.json > gpurun_out/r2c53_lp_$ns.txt 2>&1
  echo "== ns $ns"; grep -E "^(lcblock3|lcblock5|pwconv_tc|se_pwconv|total)" gpurun_out/r2c53_lp_$ns.txt | awk '{a[$1]+=$5} END {for (k in a) print k, a[k]}'
done
grep "fused\]" gpurun_out/r2c53_lp_0.txt | sort | uniq -c | sort -rn | head -40
