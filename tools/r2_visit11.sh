#!/bin/bash
# Round-2 visit 11 (ONE GPU, ~2 min): small matrices (AED-window sizes): grid of the persistent panel kernel (fewer CTAs = cheaper
# grid barriers; the GEMV of such sizes is not bandwidth-bound) and panel width.
mkdir -p gpurun_out
: > gpurun_out/sweep_small_ctas.log
for n in 1000 2000 4000 8000; do
    timeout 100 python tools/sweep.py $n "" "FUSED_CTAS=32" "FUSED_CTAS=64" "FUSED_CTAS=96" "FUSED_CTAS=128" "AUTO_PANEL_WIDTH=128" "AUTO_PANEL_WIDTH=128,FUSED_CTAS=64" 2>&1 | tee -a gpurun_out/sweep_small_ctas.log
done
