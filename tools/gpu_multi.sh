#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): real NVLink peers. usage: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/multi_smi.txt
nvidia-smi topo -m >> gpurun_out/multi_smi.txt 2>&1
# threads of one process (what starneig_node_init(cores, N, ..) + starneig_SEP_SM_Hessenberg do), parity vs the oracle
(timeout 120 python tools/multi_check.py $N 3000 2>&1 | tail -4) | tee gpurun_out/multi_threads.log
grep -q "^form 0" gpurun_out/multi_threads.log || { echo "retry without staging overlap"; (STARNEIG_B200_STAGE_OVERLAP=0 timeout 120 python tools/multi_check.py $N 3000 2>&1 | tail -4) | tee -a gpurun_out/multi_threads.log; }
(timeout 150 python tools/multi_check.py $N 6000 -1 2 2>&1 | tail -4) | tee -a gpurun_out/multi_threads.log
# one process per GPU: the bench line at N GPUs
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
(timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "processes_torchrun or match_single or col_block or partial" 2>&1 | tail -3) | tee gpurun_out/pytest_multi.log
