#!/bin/bash
# One GPU-box visit: overlap sweep, parity tests, bench line, ncu launch list of the bench command and full
# captures of the dominant kernels. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
: > gpurun_out/sweep.log
for cfg in "OVERLAP=0" "" "SIDE_FAT=0" "OVERLAP_CTAS=132" "OVERLAP_CTAS=116"; do
    timeout 120 python tools/sweep.py 20000 "$cfg" 2>&1 | grep -v "zeros below" | tee -a gpurun_out/sweep.log
done
timeout 480 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_n20000.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_panel_fused --launch-skip 2 -c 1 -o gpurun_out/panel_fused_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_fused.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel --launch-skip 8 -c 8 -o gpurun_out/dgemm_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_dgemm.log 2>&1
ls -la gpurun_out
