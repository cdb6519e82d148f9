"""bench.py on a machine without a GPU: the reference arm (CPU oracle port) prints a contract-conforming JSON line, and
our own arm refuses to run instead of falling back to anything."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_a_conforming_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"].startswith("views/sec fwd+bwd") and d["unit"] == "views/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference-cuda") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "diff_gauss_pose" in d["config"]["note"]


def test_our_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is visible")
    r = _run("--steps", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
