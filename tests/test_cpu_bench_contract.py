"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the contract's keys, runs the oracle
port on a bounded prefix and says how large the prefix was (round 1 reported the full workload's size there)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", *extra],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line_pipeline():
    d = _run("--ref-sample", "150000")
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["points_sampled"] == 150000 and cb["extrapolated"] is True
    assert set(cb["stage_seconds_per_pass"]) == {"voxel", "ground", "dbscan", "boxes"}
    assert d["config"]["points_sampled"] == 150000 and d["config"]["same_config"] is False
    assert d["config"]["full_workload_points_per_gpu"] == 100_000_000


def test_reference_arm_line_corridor_geo():
    d = _run("--workload", "corridor1B_geo", "--ref-sample", "120000")
    assert d["scaling"] == "strong" and d["config"]["ground"] == "grid"
    assert set(d["cpu_baseline"]["stage_seconds_per_pass"]) == {"voxel", "ground", "dbscan", "boxes", "geoid_crs"}
