"""CPU: bench.py's reference arm (the unmodified reference's OpenMP CCD++ on the host cores, oracle/_ref) on the small
BASELINE configs[0] shape, and the keys of the JSON line the driver reads.  The B200 arm needs a GPU and is exercised on
the GPU box; it must refuse to run without one (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=600)


def test_reference_arm_line(ref):
    r = _bench("--impl", "reference", "--workload", "ml100k_k10", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "ccdpp_seconds_per_outer_iteration" and line["unit"] == "s"
    assert line["higher_is_better"] is False and line["value"] > 0 and line["ms_per_step"] == pytest.approx(line["value"] * 1e3)
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("CCD++ k=10") and "model" not in line["config"]


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _bench("--workload", "ml100k_k10", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
