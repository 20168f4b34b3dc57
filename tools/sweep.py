"""Development sweep over engine knobs (environment variables read when the rank opens) at one matrix size,
device-resident arm. usage: sweep.py n "K1=V1,K2=V2" "K1=V3" ...   (an empty string is the default configuration)
Prints one line per configuration and checks every result against the first one (to rounding)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import starneig_b200 as sn
from bench import ClockSampler

n = int(sys.argv[1])
configs = sys.argv[2:] or [""]
ld = (n + 15) // 16 * 16
g = torch.Generator(device="cuda").manual_seed(1)
A0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=g)
A = torch.empty_like(A0)
Q = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
ref = None
KNOBS = ["STARNEIG_B200_OVERLAP", "STARNEIG_B200_OVERLAP_CTAS", "STARNEIG_B200_FUSED_CTAS", "STARNEIG_B200_SIDE_CHUNK",
         "STARNEIG_B200_FUSED_PANEL", "STARNEIG_B200_SIDE_FAT", "STARNEIG_B200_SIDE_RATE", "STARNEIG_B200_SIDE_MAX_SMS"]
for cfg in configs:
    for k in KNOBS:
        os.environ.pop(k, None)
    reps = 1
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=")
        if k == "REPS":
            reps = int(v)
        else:
            os.environ["STARNEIG_B200_" + k] = v
    sn.starneig_node_init(-1, 1, sn.STARNEIG_NO_MESSAGES)
    sn.set_profile_level(1)
    for it in range(reps):
        sampler = ClockSampler(0); sampler.start()
        A.copy_(A0)
        Q.zero_(); Q[:, :n].fill_diagonal_(1.0)
        torch.cuda.synchronize()
        r = sn.hessenberg_device(n, A, ld, Q, ld)
        torch.cuda.synchronize()
        st = sn.get_stats()
        ck = sampler.stop()
        gbs = st["gemv_timed_bytes"] / max(st["gemv_ms"], 1e-9) / 1e6
        print(f"[{cfg or 'default':40s}] ret {r} device_ms {st['device_ms']:8.1f} GFLOP/s {10 / 3 * n ** 3 / st['device_ms'] / 1e6:7.0f} "
              f"col {st['panel_ms']:7.1f} trail {st['trail_ms']:6.1f} deferred {st['other_ms']:7.1f} tail {st['side_tail_ms']:6.1f} "
              f"gemv_ms {st['gemv_ms']:7.1f} ({gbs:5.0f} GB/s) ph {[round(x) for x in st['fused_phase_ms']]} ovl {st['overlap']} sm_mhz {ck['sm_mhz']} pwr_max {ck['power_w_max']} {ck['reasons']}", flush=True)
    sn.starneig_node_finalize()
    if ref is None:
        ref = (A.clone(), Q.clone())
        below = float(torch.tril(A[:, :n].T, diagonal=-2).abs().max()) if n <= 12000 else float(torch.tril(A[:4096, :4096].T, diagonal=-2).abs().max())
        print("   zeros below the sub-diagonal:", below == 0.0, flush=True)
    else:
        dA = float((A - ref[0]).abs().max()) / float(ref[0].abs().max())
        dQ = float((Q - ref[1]).abs().max())
        print(f"   vs first config: |dH|/max|H| {dA:.2e}  |dQ| {dQ:.2e}", flush=True)
