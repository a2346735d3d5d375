"""CPU: the reference arm of bench.py (`--impl reference`) runs the reference's own binaries on the WHOLE named workload,
prints one JSON line with the contract's keys, and bounds itself by wall time, never by a smaller sample."""
import json
import os
import subprocess
import sys

import pytest

import refrun as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not R.have_ref(100, 1), reason="oracle/_ref not built")


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "30000", "--genome", "150000"] + list(extra),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [l for l in r.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--steps", "2", "--warmup", "1")
    assert d["impl"] == "reference" and d["unit"] == "Mreads/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "30000 x 100bp" in d["config"]["workload"] and "full workload" in d["config"]["sample"]
    assert abs(d["value"] - 30000 / (d["ms_per_step"] / 1000.0) / 1e6) < 1e-9


def test_reference_arm_stops_at_its_wall_time_budget_but_never_samples():
    d = _run("--steps", "50", "--warmup", "1", "--ref-budget-s", "3")
    assert 1 <= d["steps"] < 50 and d["steps_requested"] == 50      # fewer timed steps, each over the whole workload
    assert "30000 x 100bp" in d["config"]["workload"]
