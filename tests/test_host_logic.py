"""CPU-only tests: the C-ABI library loads and exports every declared symbol; the link planner."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from bendy2d_b200 import _lib, plan_links, scenes
from bendy2d_b200.solver import LinkPanic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bendy2d_b200.h")).read()
    declared = set(re.findall(r"\b(bendy_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"bendy_solver", "bendy_schedule_info"}
    assert len(declared) >= 40
    L = C.CDLL(_lib.build())
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/bendy2d_b200.h but not exported"
    assert declared == set(_lib.signatures()), "ctypes table and header out of sync"
    assert _lib.lib().bendy_abi_version() == 1


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bendy2d_b200 import BendyError, Solver

    with pytest.raises(BendyError) as e:
        Solver()
    assert "no CPU path" in str(e.value)


def _check_plan(n, ab, pack=0, maxp=0):
    rank, perm, colour, part, info = plan_links(n, ab, pack, maxp)
    ab = np.asarray(ab).reshape(-1, 2)
    m = len(ab)
    assert sorted(rank.tolist()) == list(range(n))
    assert sorted(perm.tolist()) == list(range(m))
    # links of one (partition, colour) bucket share no vertex
    key = part.astype(np.int64) * 4096 + colour
    for k in np.unique(key):
        sel = ab[key == k].ravel()
        assert len(np.unique(sel)) == len(sel), "colour bucket is not an independent set"
    # local links stay inside one partition range of at most max_points points
    # perm is bucket-major: partition-major, colour-major, global colours last
    pk = key[perm]
    is_glob = part[perm] == 0xFFFFFFFF
    assert (np.diff(is_glob.astype(int)) >= 0).all()
    loc = pk[~is_glob]
    assert (np.diff(loc) >= 0).all()
    assert (np.diff(colour[perm][is_glob].astype(np.int64)) >= 0).all()
    assert info["n_local_links"] + info["n_global_links"] == m
    return rank, perm, colour, part, info


def test_planner_lattice_body_is_one_partition():
    pos, ab = scenes.lattice_body(20, 25, 0.25, (0, 0), False)
    assert len(pos) == 500 and len(ab) == 1411
    rank, perm, colour, part, info = _check_plan(500, ab)
    assert info["n_partitions"] == 1 and info["n_global_links"] == 0
    assert info["n_local_colours"] <= 11  # greedy bound 2*deg-1 with deg 6


def test_planner_c1_link_count_and_colours():
    sc = scenes.c1_softbody_blob()
    assert sc.n_particles == 400 and sc.n_links == 1482
    assert (sc.links_ab[:, 0] < sc.links_ab[:, 1]).all()
    _check_plan(400, sc.links_ab)


def test_planner_many_bodies_pack_and_renumber():
    sc = scenes.c3_softbody_field(4, 3, 0, 0)
    assert sc.n_particles == 6000 and sc.n_links == 12 * 1411
    rank, perm, colour, part, info = _check_plan(sc.n_particles, sc.links_ab)
    assert info["n_global_links"] == 0
    assert info["n_partitions"] == 12  # pack target 512: one 500-point body each
    rank, perm, colour, part, info = _check_plan(sc.n_particles, sc.links_ab, pack=1024, maxp=4096)
    assert info["n_partitions"] == 6


def test_planner_giant_component_gets_global_colours():
    pos, ab = scenes.lattice_body(100, 100, 1.0, (0, 0), True)
    rank, perm, colour, part, info = _check_plan(10000, ab, pack=512, maxp=1024)
    assert info["n_partitions"] >= 10
    assert info["n_global_links"] > 0 and info["n_global_colours"] >= 1


def test_planner_unlinked_points_go_last_and_random_graph():
    rng = np.random.default_rng(7)
    n = 3000
    a = rng.integers(0, n - 1, 5000)
    b = rng.integers(0, n, 5000)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    keep = lo < hi
    ab = np.stack([lo[keep], hi[keep]], 1)
    rank, perm, colour, part, info = _check_plan(n, ab, pack=256, maxp=512)
    linked = np.zeros(n, bool)
    linked[ab.ravel()] = True
    if (~linked).any():
        assert rank[~linked].min() > rank[linked].max()


def test_planner_rejects_bad_links():
    with pytest.raises(LinkPanic):
        plan_links(4, [[1, 1]])
    with pytest.raises(LinkPanic):
        plan_links(4, [[0, 4]])


def test_planner_duplicate_links_get_distinct_colours():
    rank, perm, colour, part, info = _check_plan(3, [[0, 1], [0, 1], [1, 2]])
    assert colour[0] != colour[1]


def test_scene_sizes_match_survey():
    assert scenes.c2_free_particles().n_particles == 100_000
    sc = scenes.c3_softbody_field(5, 2, 4, 3)
    assert sc.n_particles == 5000 and len(sc.circles_r) == 4 and len(sc.polygons) == 3
    b = sc.algorithmic_bytes()
    assert b["K3_links"] == (sc.n_links + 18) * 44 and b["K1_integrate"] == sc.n_points * 32


def test_header_is_plain_c_and_every_entry_point_links(tmp_path):
    """gcc -std=c99 -Wall -Werror on a C program with one typed call site per entry point."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    libdir = os.path.dirname(_lib.build())
    exe = str(tmp_path / "c_abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_check.c"), "-o", exe, "-L" + libdir, "-lbendy2d_b200",
                           "-Wl,-rpath," + libdir])
    src = open(os.path.join(ROOT, "tests", "c_abi_check.c")).read()
    hdr = open(os.path.join(ROOT, "include", "bendy2d_b200.h")).read()
    declared = set(re.findall(r"\b(bendy_[a-z0-9_]+)\s*\(", hdr)) - {"bendy_solver", "bendy_schedule_info"}
    missing = [n for n in sorted(declared) if n + "(" not in src]
    assert not missing, f"c_abi_check.c does not call {missing}"
    # runs anywhere: without a GPU it stops after bendy_create fails loudly
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr


def test_rust_sys_crate_declares_the_header_symbol_for_symbol():
    """The Rust -sys crate cannot be compiled in this image (no cargo/rustc), so at least keep its extern
    block in step with include/bendy2d_b200.h: same symbol set, same number of parameters per function."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "bendy2d_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    rust = open(os.path.join(root, "rust", "bendy2d-sys", "src", "lib.rs")).read()
    rust = re.sub(r"//[^\n]*", "", rust)

    def arity(params):
        params = params.strip()
        return 0 if params in ("", "void") else params.count(",") + 1

    c_decl = {m.group(1): arity(m.group(2)) for m in re.finditer(r"\b(bendy_\w+)\s*\(([^()]*)\)\s*;", header)}
    r_decl = {m.group(1): arity(m.group(2)) for m in re.finditer(r"pub fn (bendy_\w+)\s*\(([^()]*)\)", rust)}
    assert len(c_decl) >= 58
    assert set(c_decl) == set(r_decl), (sorted(set(c_decl) - set(r_decl)), sorted(set(r_decl) - set(c_decl)))
    wrong = {k: (c_decl[k], r_decl[k]) for k in c_decl if c_decl[k] != r_decl[k]}
    assert not wrong, wrong


def test_the_emulation_build_is_not_accepted_as_a_product_library(tmp_path):
    """tests/cuemu's library only loads when the test harness opts in (BENDY_CUDA_EMU=1): a stray
    BENDY2D_B200_LIB can never turn it into a CPU path of the product."""
    import subprocess
    import sys

    fake = tmp_path / "libbendy2d_b200_emu.so"
    fake.write_bytes(b"")
    env = dict(os.environ, BENDY2D_B200_LIB=str(fake))
    env.pop("BENDY_CUDA_EMU", None)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", "from bendy2d_b200 import _lib; _lib.lib()"], cwd=root, env=env,
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
