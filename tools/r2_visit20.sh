#!/bin/bash
# Round-2 visit 20 on 2 real GPUs (charged 2x, ~4 min): backward accumulation of Q on several ranks: the multi-GPU parity tests on
# real NVLink peers, the bench line at 2 GPUs (forward order for comparison first: device arm only).
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s 2>&1 | grep -v "^\*\|OMP_NUM" | tail -8) | tee gpurun_out/pytest_multi_2gpu.log
STARNEIG_B200_Q_BACKWARD=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_n2_qforward.json 2> gpurun_out/bench_n2_qforward.err
echo "bench (forward) exit $?"; cut -c1-700 gpurun_out/bench_n2_qforward.json; tail -3 gpurun_out/bench_n2_qforward.err
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench exit $?"; cat gpurun_out/bench_n2.json; grep -v "^\*\|OMP_NUM\|^$\|ProcessGroupNCCL" gpurun_out/bench_n2.err | tail -4
