#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list (first panel of n=20000) and full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 2 --warmup 1 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_n20000_panel0.csv python tools/run_once.py 20000 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_col_gemv --launch-skip 150 -c 1 -o gpurun_out/gemv_full -f python tools/run_once.py 20000 > gpurun_out/ncu_gemv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel -c 5 -o gpurun_out/dgemm_full -f python tools/run_once.py 20000 > gpurun_out/ncu_dgemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_col_finish_update\|k_col_reflector --launch-skip 400 -c 2 -o gpurun_out/panel_full -f python tools/run_once.py 20000 > gpurun_out/ncu_panel.log 2>&1
ls -la gpurun_out
