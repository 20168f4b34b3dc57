#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/bin/overlap_probe 116 32 2>&1 | tee gpurun_out/overlap_probe.log
timeout 120 tools/bin/overlap_probe 132 16 2>&1 | tee -a gpurun_out/overlap_probe.log
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_hessenberg.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 200 python tools/sweep.py 20000 "OVERLAP=0" "OVERLAP=1,OVERLAP_CTAS=128" 2>&1 | tee gpurun_out/sweep.log
