"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference
arm prints exactly one JSON line with the agreed keys (and nothing else on stdout), and our arm
refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

import refimpl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    if not (refimpl.have_ref() or os.path.exists(os.path.join(ROOT, "oracle", "liboracle_port.so"))):
        pytest.skip("no oracle built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nodes", "2000",
                        "--steps", "1", "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "interp queries/sec" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["nodes"] == 2000 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nodes", "2000",
                        "--steps", "1", "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=120,
                       env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1",
                        "--nodes", "1000", "--queries", "1000"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert "no CPU path" in r.stderr and r.stdout.strip() == ""
