#!/bin/bash
# Round-2 last visit on ONE GPU (a few GPU-minutes were left): the reference arm of HEAD on the box's host cores (both schedules of
# the reference's task graph, oracle/ref_shim/mini_starpu.c) and a full ncu capture of the TMA-fed DMMA kernels at the automatic
# panel width 192 (the earlier captures were taken at width 312).
mkdir -p gpurun_out
timeout 170 python bench.py --impl reference --steps 2 --warmup 2 > gpurun_out/bench_reference_final3.json 2> gpurun_out/bench_reference_final3.err
echo "reference arm exit $?"; cat gpurun_out/bench_reference_final3.json; tail -3 gpurun_out/bench_reference_final3.err
timeout 150 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma_kernel --launch-skip 24 -c 8 -o gpurun_out/dgemm_tma_w192_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_dgemm_w192.log 2>&1; echo "ncu dgemm exit $?"; tail -3 gpurun_out/ncu_dgemm_w192.log
