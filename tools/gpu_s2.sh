mkdir -p gpurun_out
( for cfg in "8 1 0" "8 0 0" "8 1 1" "6 1 0" "12 1 0"; do set -- $cfg; echo "threads=$1 nt=$2 spin=$3"; HSSB_HOST_THREADS=$1 HSSB_BOUNCE_NT=$2 HSSB_BOUNCE_SPIN=$3 timeout 300 python tools/pageable_sweep.py 2>&1 | tail -1; done ) | tee gpurun_out/s2_pageable.log
