#!/bin/bash
# Round-2 visit 10 (ONE GPU, ~3 min): two micro-changes of the persistent panel kernel against the build of the previous visit:
# `pre` = epilogue inputs of phase A fetched before the partial sums; `hoist` = `pre` + the first loads of every GEMV staging
# block issued before v is staged (more registers live across the staging code: spills).
mkdir -p gpurun_out
: > gpurun_out/sweep_microopts.log
for lib in "" pre hoist; do
    L=""; [ -n "$lib" ] && L="$PWD/starneig_b200/lib_exp/libstarneig_$lib.so"
    echo "=== lib ${lib:-previous}" | tee -a gpurun_out/sweep_microopts.log
    (STARNEIG_B200_LIB="$L" timeout 120 python tools/sweep.py 20000 "" 2>&1 | tail -1) | tee -a gpurun_out/sweep_microopts.log
    (STARNEIG_B200_LIB="$L" timeout 60 python tools/sweep.py 6000 "" 2>&1 | tail -1) | tee -a gpurun_out/sweep_microopts.log
done
