#!/bin/bash
# Round-2 visit 16 (ONE GPU, ~4 min): automatic panel width on one GPU revisited with the TMA-fed DMMA kernels: widths that the
# skinny tiles (96 / 104 columns) cover without padding -- 192, 208, 288, 312 -- at n = 1000 ... 30000 (visit 15: 192 beat the
# reference formula's 312 at n = 20000 by 1.5-1.9 %, 256 and 160 did not).
mkdir -p gpurun_out
: > gpurun_out/sweep_width.log
timeout 300 python tools/sweep.py 20000 "" "AUTO_PANEL_WIDTH=288" "AUTO_PANEL_WIDTH=208" "AUTO_PANEL_WIDTH=192" "AUTO_PANEL_WIDTH=192,FUSED_SLABS=0" "AUTO_PANEL_WIDTH=96" 2>&1 | tee -a gpurun_out/sweep_width.log
timeout 100 python tools/sweep.py 10000 "" "AUTO_PANEL_WIDTH=288" "AUTO_PANEL_WIDTH=208" "AUTO_PANEL_WIDTH=192" 2>&1 | tee -a gpurun_out/sweep_width.log
timeout 100 python tools/sweep.py 6000 "" "AUTO_PANEL_WIDTH=208" "AUTO_PANEL_WIDTH=192" "AUTO_PANEL_WIDTH=96" 2>&1 | tee -a gpurun_out/sweep_width.log
timeout 100 python tools/sweep.py 2000 "" "AUTO_PANEL_WIDTH=224" "AUTO_PANEL_WIDTH=192"  "AUTO_PANEL_WIDTH=96" 2>&1 | tee -a gpurun_out/sweep_width.log
timeout 100 python tools/sweep.py 1000 "" "AUTO_PANEL_WIDTH=224" "AUTO_PANEL_WIDTH=192"  "AUTO_PANEL_WIDTH=96" 2>&1 | tee -a gpurun_out/sweep_width.log
timeout 200 python tools/sweep.py 30000 "" "AUTO_PANEL_WIDTH=192" 2>&1 | tee -a gpurun_out/sweep_width.log
