#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 400 python tools/sweep.py 20000 "OVERLAP=0" "OVERLAP=0,FUSED_CTAS=132" "OVERLAP=0,FUSED_CTAS=116" "OVERLAP=0,FUSED_CTAS=100" \
   "OVERLAP=1" "OVERLAP=1,OVERLAP_CTAS=136" "OVERLAP=1,OVERLAP_CTAS=124" "OVERLAP=1,OVERLAP_CTAS=112" "OVERLAP=1,SIDE_CHUNK=4096" "OVERLAP=1,OVERLAP_CTAS=124,SIDE_CHUNK=2048" 2>&1 | tee gpurun_out/sweep.log
