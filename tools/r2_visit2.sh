#!/bin/bash
# Round-2 visit 2 (ONE GPU, ~8 min): the pruned engine with GEMV linearity and line-aligned GEMV row blocks: parity gate,
# panel-width / linearity sweep at n = 20000, the n = 10000 chain check against the saved CPU-chain eigenvalues.
mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) | tee gpurun_out/pytest_gpu.log
CFGS=("" "GEMV_LINEAR=0" "AUTO_PANEL_WIDTH=256" "AUTO_PANEL_WIDTH=288" "AUTO_PANEL_WIDTH=320" "AUTO_PANEL_WIDTH=384" "GEMV_RESIDENT_KB=40960")
timeout 300 python tools/sweep.py 20000 "${CFGS[@]}" 2>&1 | tee gpurun_out/sweep.log
(timeout 300 python tools/chain_check.py 10000 --load-cpu tests/golden/chain_n10000_cpu_eigs.npz 2>&1 | tail -6) | tee gpurun_out/chain_n10000.log
timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
