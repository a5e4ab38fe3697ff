"""bench.py output contract, checked on CPU through the reference arm (the only arm that runs without a GPU):
stdout carries exactly ONE JSON line with the keys the driver reads; under torchrun only rank 0 prints; our arm
refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(cmd, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)


def _check_reference_line(stdout):
    lines = [ln for ln in stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["higher_is_better"] is True and d["unit"] == "graphs/s" and d["value"] > 0
    assert d["vs_baseline"] is None
    assert d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    return d


def test_reference_arm_prints_one_json_line():
    r = _run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-batch", "4"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = _check_reference_line(r.stdout)
    assert d["n_gpus"] == 1 and d["steps"] == 1


def test_reference_arm_under_torchrun_only_rank0_prints():
    r = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
              "--master-port", "29533", BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
              "--cpu-batch", "4"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = _check_reference_line(r.stdout)
    assert d["n_gpus"] == 2


def test_our_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run([sys.executable, BENCH, "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())
