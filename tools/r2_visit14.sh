#!/bin/bash
# Round-2 visit 14 (ONE GPU, ~5 min): the CTA's V / VT / Y rows resident in shared memory (STARNEIG_B200_FUSED_SLABS, FusedSmem) and
# p' taken from shared memory in phase R: GPU suite first (parity), then AED-window sizes at the default width and at the AED
# client's width 224, n = 6000, n = 20000 at the default width and at the 8-GPU width 192 (level-2 phases as replicated there).
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) | tee gpurun_out/pytest_gpu.log
: > gpurun_out/sweep_slabs.log
for n in 1000 2000 4000; do
    timeout 100 python tools/sweep.py $n "FUSED_SLABS=0" "" "AUTO_PANEL_WIDTH=224,FUSED_SLABS=0" "AUTO_PANEL_WIDTH=224,FUSED_SLABS=1" "AUTO_PANEL_WIDTH=224,FUSED_SLABS=2" "AUTO_PANEL_WIDTH=224" 2>&1 | tee -a gpurun_out/sweep_slabs.log
done
timeout 100 python tools/sweep.py 6000 "FUSED_SLABS=0" "" 2>&1 | tee -a gpurun_out/sweep_slabs.log
timeout 200 python tools/sweep.py 20000 "FUSED_SLABS=0" "" "FUSED_SLABS=1" "AUTO_PANEL_WIDTH=192,FUSED_SLABS=0" "AUTO_PANEL_WIDTH=192" 2>&1 | tee -a gpurun_out/sweep_slabs.log
