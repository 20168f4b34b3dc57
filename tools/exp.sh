nvidia-smi topo -m | head -8
for it in 1 2 3 4 5 6 7 8; do timeout 120 python tools/multi_check.py 2 130 35 3 > gpurun_out/mc.log 2>&1; echo "== it=$it rc=$?"; grep -E "fatal|form" gpurun_out/mc.log | head -2; done
timeout 200 python tools/multi_check.py 2 2000 -1 2 2>&1 | tail -4
timeout 300 python tools/multi_check.py 2 10000 -1 2 2>&1 | tail -4
timeout 300 python tools/multi_check.py 1 10000 -1 2 2>&1 | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py --n 600 --pw 40 --devices 2 2>&1 | tail -8
