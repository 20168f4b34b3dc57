timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_worker.py --size 1500 --panel 100 --devices 2 2>&1 | tail -3
timeout 300 python tools/multi_check.py 2 2000 -1 2 2>&1 | tail -2
timeout 300 python tools/multi_check.py 2 20000 -1 2 2>&1 | tail -2
