#!/bin/bash
# Round-2 visit 9 (ONE GPU, ~6 min): CTA geometry of the persistent panel kernel (more GEMV warps at fewer registers: 768 x 80,
# 1024 x 64 against the default 640 x 96), and odd sizes through the invariants check.
mkdir -p gpurun_out
: > gpurun_out/sweep_geometry.log
for lib in "" t768 t768u4 t1024; do
    L=""; [ -n "$lib" ] && L="$PWD/starneig_b200/lib_exp/libstarneig_$lib.so"
    echo "=== lib ${lib:-default}" | tee -a gpurun_out/sweep_geometry.log
    (STARNEIG_B200_LIB="$L" timeout 120 python tools/sweep.py 20000 "" 2>&1 | tail -2) | tee -a gpurun_out/sweep_geometry.log
    (STARNEIG_B200_LIB="$L" timeout 60 python tools/sweep.py 6000 "" 2>&1 | tail -1) | tee -a gpurun_out/sweep_geometry.log
done
for n in 5001 7777 12345; do (timeout 120 python tools/big_check.py $n 2>&1 | tail -3) | tee -a gpurun_out/odd_sizes.log; done
