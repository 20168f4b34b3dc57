#!/bin/bash
# Round-2 final visit on ONE GPU (HEAD: automatic width 192, V slab, backward accumulation of Q): multi-rank and variant tests after
# the last change (the whole suite ran one commit earlier: profiles/r2_v20_pytest_gpu.log), smoke, the bench line, the reference arm,
# ncu launch list of the bench command and a full capture of the persistent panel kernel.
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_variants.py -m gpu -q -x 2>&1 | tail -4) | tee gpurun_out/pytest_gpu_multi_variants.log
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2) | tee gpurun_out/smoke.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_n20000.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_panel_fused --launch-skip 2 -c 1 -o gpurun_out/panel_fused_full -f \
    python tools/run_once.py 20000 > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused exit $?"
