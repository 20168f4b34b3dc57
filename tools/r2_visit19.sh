#!/bin/bash
# Round-2 visit 19 (ONE GPU, ~5 min): backward accumulation of an identity Q (engine.cuh, Rank::reduce): GPU suite (parity),
# forward vs backward at n = 2000 ... 20000, bench line (device arm + e2e through the host call) both ways.
mkdir -p gpurun_out
(timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) | tee gpurun_out/pytest_gpu.log
: > gpurun_out/sweep_qback.log
for n in 20000 10000 6000 2000; do
    timeout 200 python tools/sweep.py $n "Q_BACKWARD=0" "" 2>&1 | tee -a gpurun_out/sweep_qback.log
done
for q in 0 256; do
    STARNEIG_B200_Q_BACKWARD=$q timeout 250 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_qback_$q.json 2> gpurun_out/bench_qback_$q.err
    echo "bench Q_BACKWARD=$q exit $?"; python - <<PY
import json
d = json.load(open("gpurun_out/bench_qback_$q.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), "parity", d["parity"]["ok"], d["parity"]["residual_u"], d["parity"]["orthogonality_u"], d["parity"]["e2e_equals_device_bitwise"], "path_frac", d["roofline"]["path_frac"], d["engine"]["q_accumulation"])
PY
done
