#!/bin/bash
# Round-2 multi-GPU visit (gpurun --gpus N, N = 8 by default; charged N x the box time, so every step is bounded):
# the bench line at N GPUs (n = 20000, BASELINE configs[2]); the opt-in variants of DESIGN.md section 4.2b at N GPUs through
# the device-resident arm inside one process group (tools/dist_sweep.py); n = 50000 sharded over the N GPUs with the
# invariants evaluated on rank 0's GPU (configs[3]).
# usage: gpurun --gpus 8 --timeout 600 -- bash tools/r2_visit8.sh 8
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
run_bench() {   # label, extra env..., device arm only unless label = bench
    local label=$1; shift
    local extra="--no-e2e"; [ "$label" = "bench" ] && extra=""
    env "$@" timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus $N --steps 2 --warmup 1 --no-cpu $extra > gpurun_out/n${N}_$label.json 2> gpurun_out/n${N}_$label.err
    echo "$label exit $?"; cat gpurun_out/n${N}_$label.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/n${N}_$label.err | tail -4
}
run_bench bench STARNEIG_BENCH_N=20000
# the variants through ONE process group (tools/dist_sweep.py): ~5 s each instead of a process start-up each
BEST="GEMV_RESIDENT_KB=40960,GEMM_OPT=1"          # the winners of the n = 6000 timings
(STARNEIG_BENCH_N=20000 STARNEIG_SWEEP=";$BEST;$BEST,FUSED_LL=2;$BEST,FUSED_LL=2,COL_BLOCK=32;$BEST,FUSED_LL=2,AUTO_PANEL_WIDTH=192;$BEST,FUSED_LL=1;$BEST,GEMV_PREFETCH=32,GEMV_PREFETCH_BULK=1;GEMV_RESIDENT_KB=81920,GEMV_PREFETCH_MB=112,GEMM_OPT=1,FUSED_LL=2;$BEST,OVERLAP=2" \
    timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    tools/dist_sweep.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12) | tee gpurun_out/dist_sweep_gpus$N.log
(STARNEIG_BENCH_N=50000 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    tools/big_check_dist.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6) | tee gpurun_out/big_n50000_gpus$N.log
ls -la gpurun_out | tail -12
