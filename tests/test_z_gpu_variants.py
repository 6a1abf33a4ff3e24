"""Opt-in kernel variants (environment switches of libbendy2d_b200, all OFF by default) against the default
path and the oracle: every variant must be bit-identical to the default.

These variants were written after the round's GPU budget was spent, so they have only run on the CPU
lock-step emulation (tests/cuemu, via tests/test_emu_parity.py).  Until they have been measured on a B200
they stay out of the default `-m gpu` run: set BENDY_TEST_UNPROVEN=1 to run them on a device
(profiles/r2_variants_ab.sh does, under a timeout).
"""
import os

import numpy as np
import pytest

from bendy2d_b200 import Solver, scenes
from helpers import bits, compare_state, oracle_from_scene, sync_schedule

EMU = os.environ.get("BENDY_CUDA_EMU") == "1"
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not EMU and os.environ.get("BENDY_TEST_UNPROVEN") != "1",
                                 reason="variant not yet proven on a device: BENDY_TEST_UNPROVEN=1 runs it")]
f32 = np.float32


class env:
    """environment switches are read by bendy_create"""

    def __init__(self, **kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run(sc, n_updates, grid_cell=0.0, **switches):
    with env(**switches):
        g = Solver()
    sc.load_into(g)
    if grid_cell:
        g.set_grid_cell(grid_cell)
    for _ in range(n_updates):
        g.update(sc.dt)
    return g


def same_bits(a, b):
    for x, y in zip(a.read_particles(), b.read_particles()):
        assert np.array_equal(bits(x), bits(y))
    if a.get_circles_len():
        for x, y in zip(a.read_circles()[:2], b.read_circles()[:2]):
            assert np.array_equal(bits(x), bits(y))


def disc_scene(width, height):
    sc = scenes.c2_free_particles(60, 40)
    sc.bounds = (0.0, 0.0, float(width), float(height))
    return sc


# emulated device: 8 resident scan CTAs -> 1 tile per CTA up to 6 tiles, 2 up to 13, 4 up to 27, then the
# two-kernel scan; B200: 1184 CTAs -> 1006 / 2012 / 4024 tiles of 2048 cells
SCAN_CASES = ([(40.0, 16.0, 0.24, 1), (40.0, 32.0, 0.24, 2), (64.0, 40.0, 0.24, 4), (128.0, 64.0, 0.24, 1)] if EMU else
              [(128.0, 128.0, 0.42, 1), (768.0, 700.0, 0.42, 2), (1024.0, 800.0, 0.42, 4), (2048.0, 2048.0, 0.42, 1)])


@pytest.mark.parametrize("width,height,cell,tiles_per_cta", SCAN_CASES)
def test_multi_tile_fused_scan_matches_default_and_oracle(width, height, cell, tiles_per_cta):
    sc = disc_scene(width, height)
    n = 40 if EMU else 120
    a = run(sc, n, grid_cell=cell)
    b = run(sc, n, grid_cell=cell, BENDY_SCAN_MT=1)
    used = b.stats()["scan_tiles_per_cta"]
    # on a device the resident-CTA capacity comes from the occupancy query: only the emulated one is known here
    assert used == tiles_per_cta if EMU else used in (1, 2, 4), used
    same_bits(a, b)
    c = run(sc, n, grid_cell=cell, BENDY_SCAN_FUSED=0)
    same_bits(a, c)
    # and against the oracle fed the same grid
    g = run(sc, 0, grid_cell=cell, BENDY_SCAN_MT=1)
    o = oracle_from_scene(sc)
    sync_schedule(g, o, sc)
    for _ in range(n):
        g.update(sc.dt)
        o.update(sc.dt)
    compare_state(g, o, max(width, height), 1e-5, "multi-tile scan")


# ------------------------------------------------------------------------------------------------
# BENDY_NARROW_DENSE=1: lane-dense resolution of the disc contacts (k2_narrow_dense)
def test_dense_narrowphase_pile_matches_default_and_oracle():
    sc = scenes.c2_free_particles(60, 40)
    sc.bounds = (0.0, 0.0, 40.0, 16.0)  # shallow box: the pile forms within the test
    n = 150
    a = run(sc, n)
    b = run(sc, n, BENDY_NARROW_DENSE=1)
    assert b.stats()["narrow_dense"] == 1
    same_bits(a, b)
    g = run(sc, 0, BENDY_NARROW_DENSE=1)
    o = oracle_from_scene(sc)
    sync_schedule(g, o, sc)
    for k in range(n):
        g.update(sc.dt)
        o.update(sc.dt)
        if (k + 1) % 25 == 0:
            compare_state(g, o, 40.0, 1e-5, f"dense narrowphase update {k + 1}")


def test_dense_narrowphase_with_circles_polygons_and_links():
    sc = scenes.c3_softbody_field(4, 2, 6, 8)
    sc.bounds = (0.0, 0.0, 128.0, 64.0)
    sc.particles = (sc.particles - np.array([40.0, 0.0], f32)).astype(f32)
    a = run(sc, 120)
    b = run(sc, 120, BENDY_NARROW_DENSE=1)
    same_bits(a, b)
    sc4 = scenes.c4_polygon_heavy(6, 150)
    same_bits(run(sc4, 100), run(sc4, 100, BENDY_NARROW_DENSE=1))


def test_dense_narrowphase_pool_flushes_when_a_warp_collects_more_pairs_than_it_holds():
    # 40 x 24 discs squeezed into a patch 12 disc-diameters wide: every disc overlaps dozens of others, so
    # a warp gathers far more than NARROW_POOL (192) pairs and has to resolve in several rounds; plus a
    # non-finite disc and discs outside the bounds (clamped into the border cells)
    rng = np.random.default_rng(7)
    pts = (np.array([10.0, 10.0]) + rng.uniform(0.0, 2.4, size=(960, 2))).astype(f32)
    pts[17] = [np.nan, 3.0]
    pts[400] = [-5.0, 10.5]
    pts[401] = [-5.05, 10.45]
    sc = scenes.Scene("crowd", (0.0, 0.0, 32.0, 32.0), particle_radius=0.1, particles=pts)
    for cell in (0.0, 0.2, 0.42):
        same_bits(run(sc, 12, grid_cell=cell), run(sc, 12, grid_cell=cell, BENDY_NARROW_DENSE=1))


def test_dense_narrowphase_on_strips_matches_single_solver():
    from bendy2d_b200 import strips
    from test_gpu_strips import touching_field

    sc = touching_field()
    ref = run(sc, 0)
    ref.update(sc.dt, n=60)
    with env(BENDY_NARROW_DENSE=1):
        grp = strips.LocalStripGroup(sc, 3)
    grp.update(sc.dt, n=60)
    assert all(o == 0 and st == 0 for _, _, o, st in grp.halo_stats())
    pos, prev = grp.read_particles()
    rp, rq = ref.read_particles()
    assert np.array_equal(bits(pos), bits(rp)) and np.array_equal(bits(prev), bits(rq))


# ------------------------------------------------------------------------------------------------
# BENDY_HALO_FUSED=1: send-buffer reset + ghost histogram in one launch; and everything switched on together
@pytest.mark.parametrize("switches", [{"BENDY_HALO_FUSED": 1},
                                      {"BENDY_HALO_FUSED": 1, "BENDY_NARROW_DENSE": 1, "BENDY_SCAN_MT": 1}])
def test_strip_variants_match_single_solver(switches):
    from bendy2d_b200 import strips
    from test_gpu_strips import touching_field

    sc = touching_field()
    ref = run(sc, 0)
    with env(**switches):
        grp = strips.LocalStripGroup(sc, 4)
    for k in range(3):
        ref.update(sc.dt, n=20)
        grp.update(sc.dt, n=20)
        pos, prev = grp.read_particles()
        rp, rq = ref.read_particles()
        assert np.array_equal(bits(pos), bits(rp)) and np.array_equal(bits(prev), bits(rq)), f"after {20 * (k + 1)} substeps"
    stats = grp.halo_stats()
    assert all(o == 0 and st == 0 for _, _, o, st in stats), stats
    assert sum(a + b for a, b, _, _ in stats) > 0


# ------------------------------------------------------------------------------------------------
# BENDY_SCATTER_ILP=1: 4 discs per thread in the counting-sort scatter
def test_scatter_with_four_discs_per_thread_matches_default():
    sc = scenes.c2_free_particles(60, 40)
    sc.bounds = (0.0, 0.0, 40.0, 16.0)
    same_bits(run(sc, 120), run(sc, 120, BENDY_SCATTER_ILP=1))
    sc3 = scenes.c3_softbody_field(4, 2, 6, 8)
    sc3.bounds = (0.0, 0.0, 128.0, 64.0)
    sc3.particles = (sc3.particles - np.array([40.0, 0.0], f32)).astype(f32)
    same_bits(run(sc3, 80), run(sc3, 80, BENDY_SCATTER_ILP=1, BENDY_NARROW_DENSE=1))
    # inverse masses need the sorted ids; a particle count that is not a multiple of the 1024 discs per CTA
    rng = np.random.default_rng(3)
    pts = (np.array([5.0, 5.0]) + rng.uniform(0.0, 6.0, size=(1531, 2))).astype(f32)
    pts[11] = [np.nan, 1.0]
    scc = scenes.Scene("crowd", (0.0, 0.0, 32.0, 32.0), particle_radius=0.1, particles=pts)
    k = rng.choice(np.array([0.0, 0.5, 1.0, 2.0], f32), len(pts)).astype(f32)
    a, b = run(scc, 0), run(scc, 0, BENDY_SCATTER_ILP=1)
    for g in (a, b):
        g.set_particle_inv_mass(k)
        g.update(scc.dt, n=25)
    same_bits(a, b)


# ------------------------------------------------------------------------------------------------
# BENDY_SORT_FUSED=1: scan + scatter in one launch (second grid barrier, generation-counted barrier words)
@pytest.mark.parametrize("width,height,cell,fits", [(40.0, 16.0, 0.24, True), (128.0, 64.0, 0.24, False)] if EMU else
                         [(128.0, 128.0, 0.42, True), (2048.0, 2048.0, 0.42, False)])
def test_fused_scan_scatter_matches_default(width, height, cell, fits):
    sc = disc_scene(width, height)
    n = 60 if EMU else 150
    a = run(sc, n, grid_cell=cell)
    b = run(sc, n, grid_cell=cell, BENDY_SORT_FUSED=1)
    same_bits(a, b)
    assert b.stats()["sort_fused_capacity"] > 0
    ka, kb = a.schedule_info()["kernels_per_substep"], b.schedule_info()["kernels_per_substep"]
    assert kb == (ka - 1 if fits else ka), (ka, kb)
    # with inverse masses (sorted ids) and through many graph replays of a multi-substep update
    sc2 = disc_scene(width, height)
    sc2.sub_steps, sc2.dt = 4, float(f32(4 / 120.0))
    k = np.where(np.arange(sc2.n_particles) % 7 == 0, 0.25, 1.0).astype(f32)
    c, d = run(sc2, 0, grid_cell=cell), run(sc2, 0, grid_cell=cell, BENDY_SORT_FUSED=1, BENDY_NARROW_DENSE=1)
    for g in (c, d):
        g.set_particle_inv_mass(k)
        g.update(sc2.dt, n=20)
    same_bits(c, d)
