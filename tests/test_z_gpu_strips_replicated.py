"""Strips with replicated polygons and circles, and with inverse masses (bit-identical to the unsharded run; the
replicated-circles all-reduce and the NCCL path ran on 2 x B200 in round 2: profiles/r2_strips_2gpu_*.txt)."""
import numpy as np
import pytest

from bendy2d_b200 import Solver, scenes, strips
from helpers import bits, max_ulp
from test_gpu_strips import touching_field

pytestmark = pytest.mark.gpu
f32 = np.float32


def test_strips_with_replicated_polygons_match_single_solver():
    """Polygons are replicated on every strip (nothing a particle does reaches a polygon): bodies rain on a row of
    static obstacles and on a dynamic overlapping polygon pair across 3 strips; the particles AND every strip's
    copy of the polygons must equal the unsharded run bit for bit."""
    sc = scenes.c3_softbody_field(8, 2, 0, 12)
    col = (np.arange(sc.n_particles) // 500) % 8
    sc.particles[:, 0] -= (col * 3.1).astype(f32)
    sc.bounds = (0.0, 0.0, 128.0, 64.0)
    # the obstacles right under the bodies (x 56..95, y 8..23); two of them dynamic and overlapping
    placed = []
    for j, p in enumerate(sc.polygons):
        placed.append((p - p.mean(0) + np.array([57.0 + 3.3 * j, 27.5])).astype(f32))
    placed[1] = (placed[0] + np.array([1.0, 0.5], f32)).astype(f32)
    sc.polygons = placed
    sc.polygons_static = [False, False] + [True] * (len(placed) - 2)
    ref = Solver()
    sc.load_into(ref)
    grp = strips.LocalStripGroup(sc, 3)
    checked = 0
    for k in range(5):
        ref.update(sc.dt, n=20)
        grp.update(sc.dt, n=20)
        if any(o or st for _, _, o, st in grp.halo_stats()):
            break  # a body bounced further sideways than the stray margin: from here on the run needs a rebalance
        rp, rq = ref.read_particles()
        gp, gq = grp.read_particles()
        assert max_ulp(gp, rp) == 0 and max_ulp(gq, rq) == 0, f"after {20 * (k + 1)} substeps"
        for sv in grp.solvers:
            for j in range(len(sc.polygons)):
                a, b = sv.read_polygon(j), ref.read_polygon(j)
                assert max_ulp(a[0], b[0]) == 0 and max_ulp(a[1], b[1]) == 0 and max_ulp(a[2], b[2]) == 0, (k, j)
        checked = k + 1
    assert checked >= 2, "the halo went stale before the bodies reached the obstacles"
    free = Solver()  # the obstacles really were hit: without them the particles end up elsewhere
    sc2 = scenes.Scene(sc.name, sc.bounds, particle_radius=sc.particle_radius, particles=sc.particles,
                       links_ab=sc.links_ab, links_len=sc.links_len)
    sc2.load_into(free)
    free.update(sc.dt, n=20 * checked)
    assert not np.array_equal(bits(free.read_particles()[0]), bits(rp))


def test_strips_with_replicated_circles_match_single_solver():
    """Circles are replicated; the fixed-point corrections each strip's own discs collect for a Circle are summed
    over the strips before the Circles' tail applies them (integer sums: same bits as the unsharded sum).  Bodies
    fall onto a row of Circles (two of them linked, two overlapping) across 3 strips."""
    sc = scenes.c3_softbody_field(8, 2, 0, 0)
    col = (np.arange(sc.n_particles) // 500) % 8
    sc.particles[:, 0] -= (col * 3.1).astype(f32)
    sc.bounds = (0.0, 0.0, 128.0, 64.0)
    sc.circles_pos = np.stack([57.0 + 3.9 * np.arange(10), np.full(10, 26.0)], 1).astype(f32)
    sc.circles_pos[3] = sc.circles_pos[2] + np.array([0.8, 0.3], f32)  # an overlapping pair: the exact circle pass runs
    sc.circles_r = np.array([1.2, 0.9, 1.1, 1.0, 1.4, 0.8, 1.3, 1.0, 0.7, 1.2], f32)
    ref = Solver()
    sc.load_into(ref)
    grp = strips.LocalStripGroup(sc, 3)
    checked = 0
    for k in range(5):
        ref.update(sc.dt, n=20)
        grp.update(sc.dt, n=20)
        if any(o or st for _, _, o, st in grp.halo_stats()):
            break
        rp, rq = ref.read_particles()
        gp, gq = grp.read_particles()
        assert max_ulp(gp, rp) == 0 and max_ulp(gq, rq) == 0, f"after {20 * (k + 1)} substeps"
        rc = ref.read_circles()
        for sv in grp.solvers:
            gc = sv.read_circles()
            assert max_ulp(gc[0], rc[0]) == 0 and max_ulp(gc[1], rc[1]) == 0, k
        checked = k + 1
    assert checked >= 3, "the halo went stale before the bodies had pushed the circles around"
    still = Solver()  # the particles really pushed the Circles: alone they would be somewhere else
    still.add_circles(sc.circles_pos, sc.circles_r)
    still.bounds.size[:] = (128.0, 64.0)
    still.update(sc.dt, n=20 * checked)
    assert not np.array_equal(bits(still.read_circles()[0]), bits(rc[0]))


def test_strips_with_inverse_masses_match_single_solver():
    """ext inverse masses in a sharded run: a ghost disc weighs in a contact exactly as on its owner (its scale
    travels with its position), pinned discs (k = 0) stay put on either side of an edge"""
    sc = touching_field(8, 2)
    rng = np.random.default_rng(5)
    k = rng.choice([0.25, 0.5, 1.0, 2.0, 4.0], sc.n_particles).astype(f32)
    k[rng.choice(sc.n_particles, 40, replace=False)] = 0.0  # pinned points
    ref = Solver()
    sc.load_into(ref)
    ref.set_particle_inv_mass(k)
    grp = strips.LocalStripGroup(sc, 3)
    grp.set_particle_inv_mass(k)
    for step in range(4):
        ref.update(sc.dt, n=15)
        grp.update(sc.dt, n=15)
        rp, rq = ref.read_particles()
        gp, gq = grp.read_particles()
        assert max_ulp(gp, rp) == 0 and max_ulp(gq, rq) == 0, f"after {15 * (step + 1)} substeps"
    stats = grp.halo_stats()
    assert all(o == 0 and st == 0 for _, _, o, st in stats), stats
    assert sum(a + b for a, b, _, _ in stats) > 0
    # and the weighted rule is really in play: the same run with unit masses ends elsewhere
    unit = Solver()
    sc.load_into(unit)
    unit.update(sc.dt, n=60)
    assert not np.array_equal(unit.read_particles()[0], rp)
