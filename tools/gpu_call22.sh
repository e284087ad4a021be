#!/bin/bash
# round-2 call 22: timing bisect of conv_halo_tc (which role bounds it?)
set -x
mkdir -p gpurun_out
for d in 0 1 2 4 3 7; do
  OAR_DBG_HALO=$d timeout 300 python tools/layerprof.py > gpurun_out/r2c22_lp_$d.txt 2>&1
  echo "dbg=$d"; grep -E "convkxk" gpurun_out/r2c22_lp_$d.txt | head -4
done
