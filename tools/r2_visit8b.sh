#!/bin/bash
# Round-2 multi-GPU visit b: panel width by number of GPUs (2, 4 of the box through sub-groups, 8) and the side-stream
# schedule of the deferred updates at 8 GPUs, n = 20000, ONE process group.   usage: gpurun --gpus 8 --timeout 400 -- bash tools/r2_visit8b.sh
mkdir -p gpurun_out
W="AUTO_PANEL_WIDTH"
SWEEP=";$W=96;$W=128;$W=160;$W=192;$W=224;$W=192,OVERLAP=1,OVERLAP_CTAS=132;$W=192,OVERLAP=1,OVERLAP_CTAS=124;$W=160,OVERLAP=1,OVERLAP_CTAS=128;P=4;P=4,$W=192;P=4,$W=256;P=2;P=2,$W=256"
(STARNEIG_SWEEP="$SWEEP" STARNEIG_BIG_N=0 STARNEIG_CHECK_FIRST=0 STARNEIG_T1_MS=${T1:-5062} timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port 29514 tools/visit8.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|ProcessGroupNCCL" | tail -30) | tee gpurun_out/visit8b.log
