#!/bin/bash
# Round-2 visit 17 (ONE GPU, ~3 min): feasibility of tensor work INSIDE the persistent panel kernel (deferred updates by a
# dedicated warp): one extra warp per CTA issues DMMAs on registers while the kernel runs (build flag SB_FUSED_BURN; 22 warps
# x 88 registers). `idle` = the same build with the extra warp asleep (isolates the register effect); the default library first.
mkdir -p gpurun_out
: > gpurun_out/sweep_burn.log
for lib in "" idle b64s2 b64s1 b512; do
    L=""; [ -n "$lib" ] && L="$PWD/starneig_b200/lib_exp/libstarneig_$lib.so"
    echo "=== lib ${lib:-default}" | tee -a gpurun_out/sweep_burn.log
    for n in 20000 6000; do
        (STARNEIG_B200_LIB="$L" timeout 120 python tools/sweep.py $n "" 2>&1 | tail -2) | tee -a gpurun_out/sweep_burn.log
    done
done
