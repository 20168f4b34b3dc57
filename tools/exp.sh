timeout 300 python -m pytest tests/test_gpu_hessenberg.py -x -q -m gpu -s > gpurun_out/t1.log 2>&1; tail -2 gpurun_out/t1.log
timeout 200 python tools/multi_check.py 1 2000 -1 2 2>&1 | tail -2
timeout 200 python tools/multi_check.py 1 10000 -1 2 2>&1 | tail -2
timeout 200 python tools/multi_check.py 1 20000 -1 2 2>&1 | tail -2
timeout 200 python tools/multi_check.py 2 2000 -1 2 2>&1 | tail -2
