#!/bin/bash
# round-2 call 47: the input stage is handed back only behind its data -- do the two race symptoms go away?
set -x
mkdir -p gpurun_out
OAR_CTC_GROUPS=2 timeout 600 python tools/ctc_dump_diff.py 400 2>&1 | tail -4
OAR_CTC_GROUPS=2 timeout 300 python tools/stress_determinism.py sleep 300 2>&1 | grep -E "baseline|mismatches"
OAR_REC_LANES=2 timeout 300 python tools/det_diff.py 10 2>&1 | grep -E "^run|regions"
OAR_REC_LANES=2 OAR_CTC_GROUPS=2 timeout 300 python tools/det_diff.py 10 2>&1 | grep -E "^run|regions"
OAR_REC_LANES=2 OAR_DBG_FB_OFF=8 timeout 300 python tools/det_diff.py 8 2>&1 | grep -E "^run|regions"
