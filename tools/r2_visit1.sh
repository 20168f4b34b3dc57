#!/bin/bash
# Round-2 visit 1 (ONE GPU, ~15 min of box time): parity gate, the tuning sweeps that decide the defaults (DMMA tile
# configurations incl. the interleaved-cp.async variants against cuBLAS on the exact shapes; panel width; ILV), the bench
# line, the ncu launch list of the bench command and one full capture of the dominant kernels. Everything lands in
# gpurun_out/.   usage: gpurun --timeout 1100 -- bash tools/r2_visit1.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
(STARNEIG_TEST_OPTIN=1 timeout 480 python -m pytest tests -m gpu -q 2>&1 | tail -12) | tee gpurun_out/pytest_gpu.log
(timeout 150 tools/bin/gemm_sweep 20000 2 3 2>&1) > gpurun_out/gemm_sweep_p2.txt; tail -50 gpurun_out/gemm_sweep_p2.txt
(timeout 100 tools/bin/gemm_sweep 20000 40 3 2>&1) > gpurun_out/gemm_sweep_p40.txt
: > gpurun_out/sweep.log
# all configurations in one process (tools/sweep.py); if a variant faults, the rest is repeated one process each
CFGS=("" "GEMM_OPT=1" "GEMV_RESIDENT_KB=20480" "GEMV_RESIDENT_KB=40960" "GEMV_RESIDENT_KB=81920,GEMV_PREFETCH_MB=112" "GEMV_RESIDENT_KB=40960,GEMM_OPT=1" "GEMV_KC=2048" "GEMV_RESIDENT_KB=40960,GEMM_OPT=1,GEMV_KC=2048" "FUSED_LL=2" "FUSED_LL=2,GEMV_RESIDENT_KB=40960,GEMM_OPT=1,GEMV_KC=2048" "FUSED_LL=1" "GEMV_PREFETCH=32,GEMV_PREFETCH_BULK=1" "OVERLAP=2" "OVERLAP=2,GEMV_RESIDENT_KB=40960" "FUSED_EVEN_ROWS=1" "FUSED_LL=1,FUSED_R=1" "GEMM_OPT=2" "GEMM_OPT=3" "AUTO_PANEL_WIDTH=256" "AUTO_PANEL_WIDTH=384")
timeout 600 python tools/sweep.py 20000 "${CFGS[@]}" 2>&1 | tee -a gpurun_out/sweep.log
DONE=$(grep -c "device_ms" gpurun_out/sweep.log)
if [ "$DONE" -lt "${#CFGS[@]}" ]; then
    echo "sweep stopped after $DONE configurations: running the others isolated (skipping the one that stopped it)" | tee -a gpurun_out/sweep.log
    timeout 600 python tools/sweep.py 20000 --isolate "${CFGS[@]:$((DONE + 1))}" 2>&1 | tee -a gpurun_out/sweep.log
fi
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
# the fastest correct configuration of the sweep, as a full bench line of its own (saves a second visit): the
# candidate for the new defaults
BEST=$(python tools/best_of_sweep.py gpurun_out/sweep.log)
echo "best of sweep: '${BEST}'" | tee -a gpurun_out/sweep.log
if [ -n "$BEST" ]; then
    (IFS=','; for kv in $BEST; do export "STARNEIG_B200_$kv"; done
     timeout 300 python bench.py --no-cpu > gpurun_out/bench_best.json 2> gpurun_out/bench_best.err; echo "bench (best) exit $?")
    cat gpurun_out/bench_best.json; tail -3 gpurun_out/bench_best.err
fi
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_n20000.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel --launch-skip 8 -c 8 -o gpurun_out/dgemm_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_dgemm.log 2>&1; echo "ncu dgemm exit $?"
(timeout 90 driver/bin/starneig-test --experiment hessenberg --n 10000 --seed 2019 --gpus 1 --repeat 1 --warmup 1 --hooks hessenberg residual 2>&1; echo "driver exit $?") | tee gpurun_out/driver_n10000.log
(timeout 200 python tools/chain_check.py 4000 2>&1 | tail -5) | tee gpurun_out/chain_n4000.log
ls -la gpurun_out | tail -20
