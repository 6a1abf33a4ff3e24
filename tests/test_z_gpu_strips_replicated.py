"""Strips with replicated polygons and circles (written after the round's GPU budget was spent: so far they have
only run on the CPU emulation, hence in a file that sorts behind the device-proven tests)."""
import numpy as np
import pytest

from bendy2d_b200 import Solver, scenes, strips
from helpers import bits, max_ulp

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
