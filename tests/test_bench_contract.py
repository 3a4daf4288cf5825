"""CPU tests of bench.py's contract: the reference arm prints ONE JSON line with the agreed keys (on a tiny sample),
and the B200 arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

from tests.parity import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--cpu-sample", "tiny", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DFSPH particle-updates/s" and d["unit"] == "particle-updates/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "tiny" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    r = subprocess.run([sys.executable, BENCH, "--particles", "tiny", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr and r.stdout.strip() == ""
