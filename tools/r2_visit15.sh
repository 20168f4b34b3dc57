#!/bin/bash
# Round-2 visit 15 (ONE GPU, ~3 min): does the L1 / shared-memory split bound the GEMV stream? (visit 14: every kilobyte of
# shared memory the persistent kernel takes slows its GEMV phases.) Unused shared memory added to the launch
# (FUSED_SMEM_PAD_KB), explicit carve-out, smaller fixed layout (GEMV_KC=256, narrower panel); slabs off.
mkdir -p gpurun_out
export STARNEIG_B200_FUSED_SLABS=0
timeout 400 python tools/sweep.py 20000 "" "FUSED_SMEM_PAD_KB=40" "FUSED_SMEM_PAD_KB=80" "FUSED_SMEM_PAD_KB=120" "FUSED_SMEM_PAD_KB=155" "FUSED_CARVEOUT=30" "FUSED_CARVEOUT=100" "GEMV_KC=256" "AUTO_PANEL_WIDTH=256" "AUTO_PANEL_WIDTH=256,GEMV_KC=256" "AUTO_PANEL_WIDTH=192,GEMV_KC=256" "AUTO_PANEL_WIDTH=160,GEMV_KC=256" 2>&1 | tee gpurun_out/sweep_l1_split.log
timeout 100 python tools/sweep.py 2000 "" "FUSED_SMEM_PAD_KB=40" "FUSED_SMEM_PAD_KB=100" "FUSED_SMEM_PAD_KB=180" "GEMV_KC=256" "GEMV_KC=128,AUTO_PANEL_WIDTH=128" 2>&1 | tee -a gpurun_out/sweep_l1_split.log
