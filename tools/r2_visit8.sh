#!/bin/bash
# Round-2 multi-GPU visit (gpurun --gpus N; charged N x the box time, so every step is bounded): ONE process group for the
# n = 20000 settings sweep + parity at N GPUs + n = 50000 with parity (tools/visit8.py), then the bench line at N GPUs.
# usage: gpurun --gpus 8 --timeout 600 -- bash tools/r2_visit8.sh 8
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
SWEEP=${STARNEIG_SWEEP:-";GEMV_RESIDENT_KB=65536;GEMV_RESIDENT_KB=98304,L2_BUDGET_MB=112;AUTO_PANEL_WIDTH=192;AUTO_PANEL_WIDTH=256,GEMV_RESIDENT_KB=65536;COL_BLOCK=32;GEMV_LINEAR=0;GEMM_TMA=0"}
(STARNEIG_SWEEP="$SWEEP" timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    tools/visit8.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -30) | tee gpurun_out/visit8_gpus$N.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; cat gpurun_out/bench_n$N.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n$N.err | tail -4
