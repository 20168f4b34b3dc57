#!/bin/bash
# Round-2 visit 13 (ONE GPU, ~3 min): `ysum` = the GEMV partial sums of a row fetched by several threads in one batch (phase A);
# `ysum_pfv` = that + the first V tile of phase R loaded before w2 arrives; against the build of the previous visit.
mkdir -p gpurun_out
: > gpurun_out/sweep_ysum.log
for lib in "" ysum ysum_pfv; do
    L=""; [ -n "$lib" ] && L="$PWD/starneig_b200/lib_exp/libstarneig_$lib.so"
    echo "=== lib ${lib:-previous}" | tee -a gpurun_out/sweep_ysum.log
    for n in 20000 6000 2000 1000; do
        (STARNEIG_B200_LIB="$L" timeout 120 python tools/sweep.py $n "" 2>&1 | tail -1) | tee -a gpurun_out/sweep_ysum.log
    done
done
