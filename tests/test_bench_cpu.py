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


def test_committed_bench_line_follows_the_contract():
    """profiles/r1_bench_c3_n1.json is the JSON line of `python bench.py` on a B200: every key of the bench
    contract is there and the derived figures are consistent with each other."""
    with open(os.path.join(ROOT, "profiles", "r1_bench_c3_n1.json")) as f:
        d = json.loads(f.read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == "particle-substeps/s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["dtype"] == "f32"
    assert "workload" in d["config"] and "C3" in d["config"]["workload"] and "flush" in d["config"]["l2"]
    sub = d["config"]["substeps_per_step"]
    assert abs(d["value"] - d["config"]["points"] * sub / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["alg_bytes_per_substep_kernel"] / (r["avg_launch_ms"] * 1e-3) / 1e9) / r["achieved"] < 1e-6
    whole = r["substep"]
    assert abs(whole["achieved"] - whole["alg_bytes"] * sub / (d["ms_per_step"] * 1e-3) / 1e9) / whole["achieved"] < 1e-6
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] == d["steps"] * sub * d["schedule"]["kernels_per_substep"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] == 1 and c["value"] > 0 and "sample" in c
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_round2_bench_lines_are_consistent():
    """profiles/r2_bench_c{1..4}_n1.json (final build of round 2, B200): contract keys, value = points x substeps / time,
    roofline fractions recomputable from their own fields, clocks sampled under load with no thermal / hw slowdown."""
    for w in ("c1", "c2", "c3", "c4"):
        with open(os.path.join(ROOT, "profiles", f"r2_bench_{w}_n1.json")) as f:
            d = json.loads(f.read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
            assert key in d, (w, key)
        sub = d["config"]["substeps_per_step"]
        assert abs(d["value"] - d["config"]["points"] * sub / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6, w
        r = d["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9, w
        whole = r["substep"]
        assert abs(whole["achieved"] - whole["alg_bytes"] * sub / (d["ms_per_step"] * 1e-3) / 1e9) / whole["achieved"] < 1e-6, w
        assert abs(whole["frac"] - whole["achieved"] / r["peak"]) < 1e-9, w
        assert d["gpu_launches"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
        c = d["clocks"]
        assert c["samples"] >= 3 and c["sm_mhz"] == c["sm_max_mhz"], (w, c)
        assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        assert 0 < d["e2e"]["value"] < d["value"] and d["e2e"]["h2d_bytes_per_step"] > 0
    # the headline: C3 on one GPU, contract bytes 228.4 MB per substep
    with open(os.path.join(ROOT, "profiles", "r2_bench_c3_n1.json")) as f:
        c3 = json.loads(f.read().strip().splitlines()[-1])
    assert c3["roofline"]["substep"]["alg_bytes"] == 228402400 and c3["value"] > 1.75e10
    assert c3["roofline"]["traffic"] is not None and c3["roofline"]["traffic"] < c3["roofline"]["alg_bytes_per_substep_kernel"]
