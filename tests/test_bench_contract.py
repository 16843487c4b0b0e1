"""CPU: bench.py's contract with the driver.  The reference arm runs anywhere (it times the CPU port) and must print
ONE JSON line with the agreed keys; our arm must fail loudly without a GPU (no CPU fallback)."""
import ctypes
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_cuda_driver():
    try:
        ctypes.CDLL("libcuda.so.1")
        return True
    except OSError:
        return False


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["value"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # BASELINE.json names the metric this line reports
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "images/sec" in base["metric"] and "images/sec" in d["metric"]


def test_our_arm_fails_loudly_without_a_gpu():
    if _has_cuda_driver():
        pytest.skip("a CUDA driver is present; this check is for the CPU-only container")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no number may be printed without a GPU"
    assert "fg_create" in r.stderr or "FGError" in r.stderr
