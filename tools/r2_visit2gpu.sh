#!/bin/bash
# Round-2 visit on 2 real GPUs (charged 2x): the multi-GPU parity tests on real NVLink peers, and the bench line at 2 GPUs incl.
# the one-process (thread per GPU) timing of the reference-facing call.   usage: gpurun --gpus 2 --timeout 600 -- bash tools/r2_visit2gpu.sh
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
(timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s 2>&1 | grep -v "^\*\|OMP_NUM" | tail -12) | tee gpurun_out/pytest_multi_2gpu.log
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench exit $?"; cat gpurun_out/bench_n2.json; grep -v "^\*\|OMP_NUM\|^$\|ProcessGroupNCCL" gpurun_out/bench_n2.err | tail -4
