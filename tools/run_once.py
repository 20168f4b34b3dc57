"""One device-resident reduction of a random n x n matrix (for ncu captures). usage: run_once.py n [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import starneig_b200 as sn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ld = (n + 15) // 16 * 16
gen = torch.Generator(device="cuda").manual_seed(2019)
dA0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)
dA = torch.empty_like(dA0)
dQ = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
sn.set_profile_level(1)
for _ in range(reps):
    dA.copy_(dA0); dQ.zero_(); dQ.diagonal()[:n].fill_(1.0)
    torch.cuda.synchronize()
    ret = sn.hessenberg_device(n, dA, ld, dQ, ld)
    st = sn.get_stats()
    print(f"ret {ret} device_ms {st['device_ms']:.1f} GFLOP/s {10 / 3 * n ** 3 / st['device_ms'] / 1e6:.0f} "
          f"launches {st['kernel_launches']}", flush=True)
sn.starneig_node_finalize()
