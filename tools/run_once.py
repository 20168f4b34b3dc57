"""Run one Hessenberg reduction through the C ABI (for ncu captures). usage: run_once.py n [panel_width] [device|host]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import starneig_b200 as sn
n = int(sys.argv[1]); pw = int(sys.argv[2]) if len(sys.argv) > 2 else -1
sn.starneig_node_init(-1, 1, sn.STARNEIG_NO_MESSAGES)
sn.set_profile_level(0)
ld = (n + 15) // 16 * 16
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=g)
Q = torch.zeros((n, ld), dtype=torch.float64, device="cuda"); Q[:, :n] = torch.eye(n, dtype=torch.float64, device="cuda")
r = sn.hessenberg_device(n, A, ld, Q, ld, panel_width=pw)
torch.cuda.synchronize()
st = sn.get_stats()
print("ret", r, "device_ms", st["device_ms"], "launches", st["kernel_launches"])
sn.starneig_node_finalize()
