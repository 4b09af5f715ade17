"""The bench.py contract: one JSON line on stdout with the keys the driver reads, for both arms.  Short runs (the numbers are not
judged here, only that the line is complete and self-consistent).  GPU only."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *extra], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines  # exactly one JSON line on stdout
    return json.loads(lines[0])


def test_b200_arm_line(built):
    d = _run("--steps", "40", "--warmup", "5", "--cpu-sample-frames", "3")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 40 and d["warmup"] == 5 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["value"] > 1000 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1.0  # frames/s and ms/frame agree (one stream)
    e = d["e2e"]
    assert e["value"] > 1000 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == d["unit"]
    assert d["gpu_launches"] >= 40 * 60  # > 60 kernels of this library per frame
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line(built):
    d = _run("--impl", "reference", "--steps", "3", "--warmup", "1")
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
