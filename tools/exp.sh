timeout 120 tools/bin/bar_bench 2>&1 | grep -E "^[0789]"
timeout 200 python tools/multi_check.py 1 2000 -1 2 2>&1 | tail -2
timeout 200 python tools/multi_check.py 1 20000 -1 2 2>&1 | tail -2
STARNEIG_B200_FUSED_PANEL=0 timeout 200 python tools/multi_check.py 1 20000 -1 2 2>&1 | tail -2
