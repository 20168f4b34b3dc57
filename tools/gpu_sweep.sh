#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/sweep.py 20000 "OVERLAP=0" "OVERLAP=1" "OVERLAP=1,SIDE_RATE=16" "OVERLAP=1,SIDE_RATE=30" \
   "OVERLAP=1,OVERLAP_CTAS=128" "OVERLAP=1,OVERLAP_CTAS=116" "OVERLAP=1,OVERLAP_CTAS=104" "OVERLAP=1,SIDE_RATE=22,SIDE_MAX_SMS=64" 2>&1 | tee gpurun_out/sweep.log
timeout 100 python tools/sweep.py 10000 "OVERLAP=0" "OVERLAP=1" 2>&1 | tee gpurun_out/sweep10k.log
