#!/bin/bash
# round-2 call 46: where do the CTC partials differ between runs (two epilogue groups, and the default)
set -x
mkdir -p gpurun_out
OAR_CTC_GROUPS=2 timeout 600 python tools/ctc_dump_diff.py 150 > gpurun_out/r2c46_dump2.txt 2>&1; tail -40 gpurun_out/r2c46_dump2.txt
timeout 600 python tools/ctc_dump_diff.py 300 > gpurun_out/r2c46_dump1.txt 2>&1; tail -8 gpurun_out/r2c46_dump1.txt
