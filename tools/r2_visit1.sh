#!/bin/bash
# Round-2 visit 1 (ONE GPU, ~15 min of box time): parity gate incl. the at-size tests (n = 10000, the reference ctest grid at
# n = 4000), the tuning sweeps that decide the defaults (DMMA tile configurations against cuBLAS on the exact shapes; the
# engine switches at n = 20000), the bench line with its parity object, the ncu launch list of the bench command and full
# captures of the dominant kernels. Everything lands in gpurun_out/.   usage: gpurun --timeout 1200 -- bash tools/r2_visit1.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
(timeout 420 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25) | tee gpurun_out/pytest_gpu.log
(timeout 150 tools/bin/gemm_sweep 20000 2 3 2>&1) > gpurun_out/gemm_sweep_p2.txt; tail -60 gpurun_out/gemm_sweep_p2.txt
: > gpurun_out/sweep.log
CFGS=("" "GEMM_OPT=1" "GEMV_RESIDENT_KB=40960" "GEMV_RESIDENT_KB=81920,GEMV_PREFETCH_MB=112" "GEMV_KC=2048" "GEMV_RESIDENT_KB=40960,GEMM_OPT=1,GEMV_KC=2048" "FUSED_LL=2" "FUSED_LL=2,GEMV_RESIDENT_KB=40960,GEMM_OPT=1,GEMV_KC=2048" "GEMV_PREFETCH=32,GEMV_PREFETCH_BULK=1" "OVERLAP=2" "GEMM_OPT=3" "AUTO_PANEL_WIDTH=256" "AUTO_PANEL_WIDTH=384")
timeout 420 python tools/sweep.py 20000 "${CFGS[@]}" 2>&1 | tee -a gpurun_out/sweep.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_n20000.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel --launch-skip 8 -c 6 -o gpurun_out/dgemm_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_dgemm.log 2>&1; echo "ncu dgemm exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_panel_fused --launch-skip 2 -c 1 -o gpurun_out/panel_fused_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused exit $?"
ls -la gpurun_out | tail -20
