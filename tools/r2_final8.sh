#!/bin/bash
# Round-2 final multi-GPU visit: the bench line at 8 and at 4 GPUs (device-resident arm, host arm one process per GPU, and the
# one-process reference-facing call with one host thread per GPU; parity at every arm).   usage: gpurun --gpus 8 --timeout 700 -- bash tools/r2_final8.sh
mkdir -p gpurun_out
for N in 8 4; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
        bench.py --gpus $N --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_final_n$N.json 2> gpurun_out/bench_final_n$N.err
    echo "bench N=$N exit $?"; cat gpurun_out/bench_final_n$N.json; grep -v "^\*\|OMP_NUM\|^$\|ProcessGroupNCCL" gpurun_out/bench_final_n$N.err | tail -4
done
