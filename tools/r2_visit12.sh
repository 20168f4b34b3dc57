#!/bin/bash
# Round-2 visit 12 (ONE GPU, ~5 min): extra evidence at size -- the C test-driver harness at n = 10000 (BASELINE configs[1]:
# the reference driver's own hooks on the product's result), n = 30000 and n = 50000 on ONE B200 with the invariants.
mkdir -p gpurun_out
(timeout 120 driver/bin/starneig-test --experiment hessenberg --n 10000 --seed 2019 --gpus 1 --repeat 1 --warmup 1 --hooks hessenberg residual 2>&1; echo "driver exit $?") | tee gpurun_out/driver_n10000.log
(timeout 200 python tools/big_check.py 30000 2>&1 | tail -3) | tee gpurun_out/big_n30000_1gpu.log
(timeout 400 python tools/big_check.py 50000 2>&1 | tail -3) | tee gpurun_out/big_n50000_1gpu.log
