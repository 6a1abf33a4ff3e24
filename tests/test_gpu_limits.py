"""Parity at the size limits of the single-CTA / shared-memory kernels (through the C ABI, vs the oracle):
circle counts either side of the 1024-row and 4096-circle switches of k_circles_exact, link partitions at the
shared-memory opt-in threshold, and a connected component too big for one partition (global colours)."""
import numpy as np
import pytest

from bendy2d_b200 import BendyError, Solver, scenes
from oracle import bo

from helpers import max_ulp

pytestmark = pytest.mark.gpu
f32 = np.float32


def circle_field(n, side, seed):
    """n circles, radii U(0.5, 1.5), uniformly in a side x side box: dense enough for multi-contact rows."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(2.0, side - 2.0, size=(n, 2)).astype(f32)
    rad = rng.uniform(0.5, 1.5, size=n).astype(f32)
    g, o = Solver(), bo.OracleSolver()
    g.bounds.size[:] = (side, side)
    o.set_bounds(0, 0, side, side)
    g.add_circles(pos, rad)
    for p, r in zip(pos, rad):
        o.add_circle(p, float(r))
    return g, o


@pytest.mark.parametrize("n,side,updates", [(300, 40.0, 12), (1024, 70.0, 6), (1100, 75.0, 6), (4200, 150.0, 3)])
def test_circle_pass_exact_at_every_size_class(n, side, updates):
    """<= 253 circles fit the default 48 KB; 300 and 1024 need the shared-memory opt-in of the parallel path;
    1100 takes the CTA-per-row path in shared memory; 4200 works on the global array."""
    g, o = circle_field(n, side, seed=n)
    for k in range(updates):
        g.update(1 / 120)
        o.update(1 / 120)
        gp, gq, _ = g.read_circles()
        op, oq, _ = o.circles()
        assert max_ulp(gp, op) == 0 and max_ulp(gq, oq) == 0, f"{n} circles, update {k}"
    assert np.isfinite(gp).all()


def giant_lattice(cols, rows):
    """One connected lattice; rest lengths from the regular grid, positions jittered so every link works."""
    pos, ab = scenes.lattice_body(cols, rows, 0.5, (10.0, 5.0), True)
    ln = np.sqrt(((pos[ab[:, 0]].astype(np.float64) - pos[ab[:, 1]]) ** 2).sum(1)).astype(f32)
    jitter = np.random.default_rng(cols * rows).uniform(-0.05, 0.05, size=pos.shape)
    return (pos + jitter).astype(f32), ab, ln


def run_lattice(g, o, pos, ab, ln, updates, k=None):
    g.bounds.size[:] = (128.0, 96.0)
    o.set_bounds(0, 0, 128, 96)
    g.add_particles(pos)
    g.add_particle_links(ab, ln)
    o.add_particles(pos)
    o.add_particle_links(ab, ln)
    if k is not None:
        g.set_particle_inv_mass(k)
        o.set_particle_inv_mass(0, k)
    o.set_link_order(g.link_order())
    for u in range(updates):
        g.update(1 / 120)
        o.update(1 / 120)
        assert max_ulp(g.read_particles()[0], o.particles()[0]) == 0, f"update {u}"


def test_component_larger_than_a_partition_uses_global_colours():
    pos, ab, ln = giant_lattice(70, 70)  # 4900 points in one component
    g, o = Solver(), bo.OracleSolver()
    g.set_plan_params(512, 1024)
    run_lattice(g, o, pos, ab, ln, 10)
    info = g.schedule_info()
    assert info["n_partitions"] >= 5 and info["n_global_links"] > 0 and info["n_global_colours"] >= 1


def test_partition_at_the_shared_memory_opt_in_threshold():
    """One 4096-point partition with inverse masses is 48 KB of dynamic shared memory on top of the static
    colour table: more than the default limit, so the kernel attribute must have been raised."""
    pos, ab, ln = giant_lattice(64, 64)  # exactly 4096 points, one component
    g, o = Solver(), bo.OracleSolver()
    k = np.ones(len(pos), f32)
    k[:64] = 0.0  # pin the top row
    k[2000:2100] = 0.5
    run_lattice(g, o, pos, ab, ln, 8, k)
    info = g.schedule_info()
    assert info["n_partitions"] == 1 and info["n_global_links"] == 0
    assert np.array_equal(g.read_particles()[0][:64], pos[:64])


def test_biggest_allowed_partition():
    pos, ab, ln = giant_lattice(128, 128)  # 16384 points, one component, one partition of 128 KB
    g, o = Solver(), bo.OracleSolver()
    g.set_plan_params(16384, 16384)
    run_lattice(g, o, pos, ab, ln, 4)
    assert g.schedule_info()["n_partitions"] == 1


def test_plan_params_outside_the_limits_are_rejected():
    g = Solver()
    for args in ((0, 1), (0, 16385), (16385, 0)):
        with pytest.raises(BendyError):
            g.set_plan_params(*args)
