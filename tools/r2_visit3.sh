#!/bin/bash
# Round-2 visit 3 (ONE GPU, ~8 min): the TMA-fed DMMA kernels (dgemm_tma.cuh) on hardware: tile sweep against cuBLAS and the
# cp.async kernels on the exact shapes and operand parities of the engine, parity gate with TMA as the default, engine A/B
# at n = 20000, ncu full capture.
mkdir -p gpurun_out
(timeout 300 tools/bin/gemm_sweep 20000 2 3 2>&1) > gpurun_out/gemm_sweep_tma_p2.txt; echo "gemm_sweep exit $?"; cat gpurun_out/gemm_sweep_tma_p2.txt
(timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) | tee gpurun_out/pytest_gpu.log
CFGS=("" "GEMM_TMA=0" "GEMM_TMA=1")
timeout 200 python tools/sweep.py 20000 "${CFGS[@]}" 2>&1 | tee gpurun_out/sweep.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma --launch-skip 8 -c 6 -o gpurun_out/dgemm_tma_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_dgemm.log 2>&1; echo "ncu dgemm exit $?"
