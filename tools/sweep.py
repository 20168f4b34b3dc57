"""Device-resident reductions under different STARNEIG_B200_* settings, one subprocess per configuration.
usage: sweep.py n "K1=V1,K2=V2" ["..."]   (an empty string is the default configuration)"""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, subprocess, threading
sys.path.insert(0, %r)
import torch
import starneig_b200 as sn
n = int(sys.argv[1]); label = sys.argv[2]
ld = (n + 15) // 16 * 16
gen = torch.Generator(device="cuda").manual_seed(2019)
dA0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)
dA = torch.empty_like(dA0); dQ = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
sn.set_profile_level(1)
best = None
for it in range(2):
    dA.copy_(dA0); dQ.zero_(); dQ.diagonal()[:n].fill_(1.0); torch.cuda.synchronize()
    ret = sn.hessenberg_device(n, dA, ld, dQ, ld)
    st = sn.get_stats()
    if best is None or st["device_ms"] < best["device_ms"]: best = st
assert float(torch.tril(dA[:256, :256].T, diagonal=-2).abs().max()) == 0.0, "zeros below the sub-diagonal"
st = best
gb = st["gemv_timed_bytes"] / max(st["gemv_ms"], 1e-9) / 1e6
print(f"[{label:40s}] ret {ret} device_ms {st['device_ms']:8.1f} GFLOP/s {10 / 3 * n ** 3 / st['device_ms'] / 1e6:7.0f} "
      f"col {st['panel_ms']:7.1f} trail {st['trail_ms']:6.1f} deferred {st['other_ms']:7.1f} tail {st['side_tail_ms']:6.1f} "
      f"gemv_ms {st['gemv_ms']:7.1f} ({gb:5.0f} GB/s) ph {[round(x) for x in st['fused_phase_ms']]} ovl {st['overlap']}", flush=True)
sn.starneig_node_finalize()
''' % ROOT

n = sys.argv[1]
for cfg in sys.argv[2:] or [""]:
    env = dict(os.environ)
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        env["STARNEIG_B200_" + k] = v
    subprocess.run([sys.executable, "-c", CHILD, n, cfg or "default"], env=env, check=False)
