#!/bin/bash
# Final 1-GPU visit of the round: parity tests, the bench line, the C test driver at n = 10000 (BASELINE configs[1]),
# n = 50000 on one GPU (configs[3] size), ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
(timeout 90 driver/bin/starneig-test --experiment hessenberg --n 10000 --seed 2019 --gpus 1 --repeat 1 --warmup 1 --hooks hessenberg residual 2>&1; echo "driver exit $?") | tee gpurun_out/driver_n10000.log
(timeout 220 python tools/big_check.py 50000 2>&1 | tail -6) | tee gpurun_out/big_n50000.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_n20000.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out | tail -12
