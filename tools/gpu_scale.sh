#!/bin/bash
# bench line at N GPUs (gpurun --gpus N), one process per GPU. usage: tools/gpu_scale.sh N [n]
# (the size goes through the environment: torchrun's argparse chokes on an abbreviable --n among the script's arguments)
N=${1:-8}; SIZE=${2:-20000}
mkdir -p gpurun_out
STARNEIG_BENCH_N=$SIZE timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_n${N}_${SIZE}.json 2> gpurun_out/bench_n${N}_${SIZE}.err
echo "bench exit $?"; cat gpurun_out/bench_n${N}_${SIZE}.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_${SIZE}.err | tail -8
free -g | head -2
