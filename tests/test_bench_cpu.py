"""CPU-only checks of bench.py: the reference arm runs without a GPU and prints the contract's JSON
line; the byte model matches SURVEY.md §8.4."""
import json
import os
import subprocess
import sys

from bendy2d_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-substeps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "particle-substeps/s" and d["dtype"] == "f32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "sample" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_match_survey_formula():
    sc = scenes.c3_softbody_field(5, 2, 4, 3)
    n_pts, n_disc, n_link = sc.n_points, sc.n_particles, sc.n_links + sc.n_polygon_points
    cells = sc.n_cells()
    b = sc.algorithmic_bytes()
    expect = n_pts * 32 + n_disc * 56 + cells * 16 + n_link * 44 + (
        sc.n_particles * 16 + sc.n_polygon_points * 8 + len(sc.polygons) * 24)
    assert b["total"] == expect
