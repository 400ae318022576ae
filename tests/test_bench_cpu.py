"""bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`) times the compiled
reference (oracle/_ref, else the plain-C port) on the host cores and prints ONE JSON line with the keys the driver reads;
ranks other than 0 print nothing and exit 0; the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--size", "64", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel*iter/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("c4_all_methods_64x64")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["methods"]) == {"diffusion", "curvature", "elastic", "thirion", "diffeomorphic", "fluid"}


def test_reference_arm_is_rank_zero_only():
    r = _run(["--impl", "reference", "--size", "64", "--steps", "1", "--warmup", "1"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "1", "--size", "64", "--batch", "0", "--no-fp64"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
