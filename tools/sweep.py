"""Device-resident reductions on one GPU under different STARNEIG_B200_* settings.
usage: sweep.py n "K1=V1,K2=V2" ["..."]   (keys without the STARNEIG_B200_ prefix; an empty string is the default configuration)
The settings are read when the engine is created, i.e. at the first reduction after starneig_node_init: by default all
configurations run in ONE process (node finalize / init between them; ~11 s each at n = 20000 instead of ~25 s with a fresh
interpreter, CUDA context and input matrix). --isolate runs every configuration in its own process (a faulting variant
then cannot take the rest of the sweep with it)."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_configs(n, configs):
    import torch
    import starneig_b200 as sn
    ld = (n + 15) // 16 * 16
    gen = torch.Generator(device="cuda").manual_seed(2019)
    dA0 = torch.rand((n, ld), dtype=torch.float64, device="cuda", generator=gen)
    dA = torch.empty_like(dA0)
    dQ = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    for cfg in configs:
        added = []
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            os.environ["STARNEIG_B200_" + k] = v
            added.append("STARNEIG_B200_" + k)
        sn.starneig_node_init(sn.STARNEIG_USE_ALL, 1, sn.STARNEIG_NO_MESSAGES)
        sn.set_profile_level(1)
        best, ret = None, None
        for it in range(2):
            dA.copy_(dA0); dQ.zero_(); dQ.diagonal()[:n].fill_(1.0); torch.cuda.synchronize()
            ret = sn.hessenberg_device(n, dA, ld, dQ, ld)
            st = sn.get_stats()
            if best is None or st["device_ms"] < best["device_ms"]:
                best = st
        form_ok = float(torch.tril(dA[:256, :256].T, diagonal=-2).abs().max()) == 0.0 and bool(torch.isfinite(dA[:512]).all())
        st = best
        gb = st["gemv_timed_bytes"] / max(st["gemv_ms"], 1e-9) / 1e6
        print(f"[{cfg or 'default':58s}] ret {ret} device_ms {st['device_ms']:8.1f} GFLOP/s {10 / 3 * n ** 3 / st['device_ms'] / 1e6:7.0f} "
              f"col {st['panel_ms']:7.1f} trail {st['trail_ms']:6.1f} deferred {st['other_ms']:7.1f} tail {st['side_tail_ms']:6.1f} "
              f"gemv_ms {st['gemv_ms']:7.1f} ({gb:5.0f} GB/s) ph {[round(x) for x in st['fused_phase_ms']]} ovl {st['overlap']} "
              f"slabs {st['fused_slab_panels']} form_ok {form_ok}", flush=True)
        sn.starneig_node_finalize()
        for k in added:
            os.environ.pop(k, None)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--isolate"]
    n = int(args[0])
    configs = args[1:] or [""]
    if "--isolate" in sys.argv:
        for cfg in configs:
            subprocess.run([sys.executable, os.path.abspath(__file__), str(n), cfg], check=False)
    else:
        run_configs(n, configs)
