"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`, the
reference's CPU path on the host cores) on a tiny sample, and that our arm refuses to run without a CUDA device instead
of falling back to anything."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _bench(*argv, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=e)


def test_reference_arm_line():
    r = _bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-n", "300")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] == 1
    assert j["unit"] == "GFLOP/s" and j["higher_is_better"] is True and j["dtype"] == "f64" and j["vs_baseline"] is None
    assert j["value"] > 0 and j["ms_per_step"] > 0
    # `config` is the dict our arm prints too (bench.shared_config); the bounded sample is named outside it
    assert j["config"] == {"workload": j["config"]["workload"], "n": 20000} and "n=20000" in j["config"]["workload"]
    assert j["sample_n"] == 300
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert cb["lapack"]["value"] > 0 and "dgehrd" in cb["lapack"]["what"]
    # the reference's task graph is timed under both schedules of the StarPU stand-in; the faster one carries the value
    if cb["kind"] == "reference":
        probed = cb["schedules_probed"]
        assert set(probed) == {"task-parallel", "blas-parallel"} and all(v["value"] > 0 for v in probed.values())
        assert cb["schedule"] == max(probed, key=lambda k: probed[k]["value"]) and cb["schedule"] in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    r = _bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-n", "300",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _bench("--steps", "1", "--warmup", "0", "--no-cpu", "--n", "64", env={"CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode != 0                       # no CPU fallback: the product path fails loudly
    assert not any(l.startswith("{") for l in r.stdout.splitlines())
