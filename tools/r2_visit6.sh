#!/bin/bash
# Round-2 visit 6 (ONE GPU, ~9 min): TMA-streamed GEMV in the persistent kernel vs register loads, final DMMA tiles: parity gate,
# A/B at n = 20000 and at the sizes of AED windows, bench line, ncu launch list + full captures of the kernels of HEAD.
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) | tee gpurun_out/pytest_gpu.log
timeout 200 python tools/sweep.py 20000 "" "GEMV_TMA=0" 2>&1 | tee gpurun_out/sweep.log
for n in 1000 2000 4000 6000 10000; do timeout 100 python tools/sweep.py $n "" "GEMV_TMA=0" "AUTO_PANEL_WIDTH=224" 2>&1 | tee -a gpurun_out/sweep_small.log; done
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_n20000.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:dgemm --launch-skip 8 -c 6 -o gpurun_out/dgemm_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_dgemm.log 2>&1; echo "ncu dgemm exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_panel_fused --launch-skip 2 -c 1 -o gpurun_out/panel_fused_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused exit $?"
