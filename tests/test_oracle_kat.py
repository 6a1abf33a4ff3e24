"""Known-answer tests pinning the CPU oracle.

The reference has no tests or golden vectors; these KATs are hand-derived from the cited source
lines and evaluated independently here with numpy float32 in the same operator order.
"""
import os

import numpy as np
import pytest

from oracle import bo

f32 = np.float32


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def test_k_int_1_exact_integration():
    # particle.rs:20-25, solver.rs:106-116: p=(50,50)=prev, g=(0,64), dt=0.125
    s = bo.OracleSolver()
    s.set_gravity(0.0, 64.0)
    s.add_particle(50.0, 50.0)
    s.update(0.125)
    pos, prev = s.particles()
    assert pos.tolist() == [[50.0, 51.0]] and prev.tolist() == [[50.0, 50.0]]
    s.update(0.125)
    pos, prev = s.particles()
    assert pos.tolist() == [[50.0, 53.0]] and prev.tolist() == [[50.0, 51.0]]


def test_k_int_2_bits():
    # g=(0,98.2), dt=(1/60)*0.125, three substeps -> y bit patterns from SURVEY §4
    s = bo.OracleSolver()
    s.add_particle(50.0, 50.0)
    dt = f32(f32(1.0) / f32(60.0)) * f32(0.125)
    ys = []
    for _ in range(3):
        s.update(float(dt))
        ys.append(int(bits(s.particles()[0][0, 1]).item()))
    assert ys == [0x42480070, 0x42480150, 0x424802A0]
    # independent numpy evaluation of (pos+vel)+((acc*dt)*dt)
    pos, prev, g = f32(50.0), f32(50.0), f32(98.2)
    for k in range(3):
        vel = f32(pos - prev)
        prev = pos
        pos = f32(f32(pos + vel) + f32(f32(g * dt) * dt))
        assert int(bits(pos)) == ys[k]


def test_k_int_substeps_equivalence():
    # sub_steps=8 with update(dt) == 8 x update(dt/8)  (x0.125 is exact)
    a, b = bo.OracleSolver(), bo.OracleSolver()
    for s in (a, b):
        s.add_particle(10.0, 20.0)
        s.add_particle(13.0, 24.0)
        s.add_particle_link(0, 1, 4.5)
    a.set_sub_steps(8)
    dt = float(f32(1.0) / f32(60.0))
    a.update(dt)
    for _ in range(8):
        b.update(float(f32(dt) * f32(0.125)))
    assert np.array_equal(bits(a.particles()[0]), bits(b.particles()[0]))
    assert np.array_equal(bits(a.particles()[1]), bits(b.particles()[1]))


def test_k_link_1():
    # link.rs:22-26: a=(0,0) b=(3,4) target 3
    a, b = bo.prim_link_solve([0, 0], [3, 4], 3.0)
    assert bits(a).tolist() == [0x3F19999A, 0x3F4CCCCD]
    assert bits(b).tolist() == [0x4019999A, 0x404CCCCD]
    # numpy restatement
    d = np.array([0 - 3, 0 - 4], f32)
    dist = np.sqrt(f32(d[0] * d[0]) + f32(d[1] * d[1]), dtype=f32)
    n = d / dist
    c = (n * f32(dist - f32(3.0))) * f32(0.5)
    assert np.array_equal(bits(np.array([0, 0], f32) - c), bits(a))
    assert np.array_equal(bits(np.array([3, 4], f32) + c), bits(b))


def test_k_circ_1():
    # circle.rs:33-43: c1=(0,0) r1, c2=(2,0) r2 -> c1'=(-0.8,0), c2'=(2.2,0)
    hit, p1, p2 = bo.prim_circle_solve([0, 0], [2, 0], 1.0, 2.0)
    assert hit
    assert bits(p1)[0] == 0xBF4CCCCD and p1[1] == 0.0
    assert bits(p2)[0] == 0x400CCCCD and p2[1] == 0.0
    # strict '<': touching circles do not interact
    hit, p1, p2 = bo.prim_circle_solve([0, 0], [3, 0], 1.0, 2.0)
    assert not hit and p1.tolist() == [0, 0] and p2.tolist() == [3, 0]


def test_k_circle_link():
    # link.rs:36-48 with r_a=1, r_b=2: a moves 4/5 of the error, b 1/5
    a, b = bo.prim_circle_link_solve([0, 0], [10, 0], 1.0, 2.0, 5.0)
    d = f32(-10.0)
    dist = f32(10.0)
    n = f32(d / dist)
    scale = f32(1.0) / f32(f32(1.0) + f32(4.0))
    da = f32(f32(f32(n * f32(dist - f32(5.0))) * scale) * f32(4.0))
    db = f32(f32(f32(n * f32(dist - f32(5.0))) * scale) * f32(1.0))
    assert bits(a)[0] == bits(f32(0.0) - da) and bits(b)[0] == bits(f32(10.0) + db)
    assert a[0] == pytest.approx(4.0) and b[0] == pytest.approx(9.0)


def test_k_bnd_1():
    # particle.rs:27-46: pos=(-1,101) prev=(0.5,100), bounds (0,0)+(100,100)
    pos, prev = bo.prim_particle_bounds([-1, 101], [0.5, 100], (0, 0, 100, 100))
    assert pos.tolist() == [0.0, 100.0]
    assert prev.tolist() == [-1.5, 101.0]
    # inside: untouched; exactly on the wall: untouched (strict compares)
    pos, prev = bo.prim_particle_bounds([0, 100], [3, 4], (0, 0, 100, 100))
    assert pos.tolist() == [0.0, 100.0] and prev.tolist() == [3.0, 4.0]


def test_k_circle_bounds():
    # circle.rs:11-30: radius inset on both sides
    pos, prev = bo.prim_circle_bounds([1, 99], [2, 98], 2.0, (0, 0, 100, 100))
    assert pos.tolist() == [2.0, 98.0]
    assert prev.tolist() == [1.0, 99.0]  # lo+r-(prev-pos) = 2-1 ; (100-2)-(98-99) = 99


def test_k_par_parallel_segments_none():
    # common.rs:15-18: division by zero -> all comparisons false -> None
    assert bo.prim_line_intersection([0, 0], [1, 0], [0, 1], [1, 1]) is None
    assert bo.prim_line_intersection([0, 0], [1, 0], [0, 0], [1, 0]) is None  # collinear: 0/0 = NaN


def test_line_intersection_point_on_line1():
    out = bo.prim_line_intersection([0, 0], [4, 0], [1, -1], [1, 1])
    assert out.tolist() == [1.0, 0.0]
    # end-point inclusive (s,t in [0,1])
    out = bo.prim_line_intersection([0, 0], [4, 0], [4, -1], [4, 1])
    assert out.tolist() == [4.0, 0.0]
    assert bo.prim_line_intersection([0, 0], [4, 0], [5, -1], [5, 1]) is None


def test_k_poly_1_stale_edge_last_writer_wins():
    # polygon.rs:147-216 (SURVEY §4 K-poly-1)
    A = np.array([[0, 0], [4, 0], [4, 4], [0, 4]], f32)
    B = np.array([[3, 1], [7, 1], [7, 3], [3, 3]], f32)
    A2, B2 = bo.prim_solve_polygon_single(A, [2, 2], B, [5, 2])
    np.testing.assert_allclose(A2, [[0, 0], [3.75, 0], [3.5833333, 4], [0, 4]], rtol=0, atol=1e-6)
    np.testing.assert_allclose(B2, [[4, 1], [7, 1], [7, 3], [4, 3]], rtol=0, atol=1e-6)
    # second direction: no hits
    B3, A3 = bo.prim_solve_polygon_single(B2, [5, 2], A2, [2, 2])
    assert np.array_equal(bits(A3), bits(A2)) and np.array_equal(bits(B3), bits(B2))


def test_k_nan_degenerate():
    # link.rs:23-24 / circle.rs:36-37: coincident points -> 0/0 -> NaN, no guard
    a, b = bo.prim_link_solve([1, 1], [1, 1], 1.0)
    assert np.isnan(a).all() and np.isnan(b).all()
    hit, p1, p2 = bo.prim_circle_solve([1, 1], [1, 1], 1.0, 1.0)
    assert hit and np.isnan(p1).all() and np.isnan(p2).all()


def test_k_panic_invalid_link():
    # link.rs:19-21: a >= b or b out of range panics in the reference
    for a, b in ((1, 1), (1, 0), (0, 2)):
        s = bo.OracleSolver()
        s.add_particle(0, 0)
        s.add_particle(1, 0)
        s.add_particle_link(a, b, 1.0)
        with pytest.raises(bo.OraclePanic):
            s.update(0.01)


def test_solver_defaults_and_phase_order():
    # solver.rs:34-50 defaults; :109-115 order: bounds BEFORE integrate within a substep
    s = bo.OracleSolver()
    s.add_particle(50.0, 99.99)
    dt = 0.1
    s.update(dt)  # falls: pos.y = 99.99 + 98.2*0.01 = 100.972 (not clamped this substep)
    pos, prev = s.particles()
    assert pos[0, 1] > 100.0
    s.update(dt)  # now clamped to 100, velocity reflected, then integrated off the wall
    pos, prev = s.particles()
    assert prev[0, 1] == 100.0
    y0 = f32(99.99)
    y1 = f32(f32(y0 + f32(0.0)) + f32(f32(f32(98.2) * f32(dt)) * f32(dt)))
    vel = f32(y0 - y1)  # prev - pos
    prev_ref = f32(f32(f32(0.0) + f32(100.0)) - vel)
    v2 = f32(f32(100.0) - prev_ref)
    y2 = f32(f32(f32(100.0) + v2) + f32(f32(f32(98.2) * f32(dt)) * f32(dt)))
    assert bits(pos[0, 1]) == bits(y2)


def test_static_polygon_is_not_integrated_but_still_bounded():
    # polygon.rs:125-128 (static skips integrate) and :136-140 (bounds still applied)
    s = bo.OracleSolver()
    s.add_polygon_new(np.array([[10, 10], [12, 10], [12, 12], [10, 12]], f32), True)
    s.add_polygon_new(np.array([[20, 10], [22, 10], [22, 12], [20, 12]], f32), False)
    s.update(0.01)
    ps, _, c = s.polygon(0)
    assert ps.tolist() == [[10, 10], [12, 10], [12, 12], [10, 12]]
    pd, _, _ = s.polygon(1)
    assert (pd[:, 1] > np.array([10, 10, 12, 12], f32)).all()


def test_polygon_new_links_and_center():
    # polygon.rs:84-123
    s = bo.OracleSolver()
    s.add_polygon_new(np.array([[0, 0], [3, 0], [3, 4]], f32), False)
    ab, ln = s.polygon_links(0)
    assert ab.tolist() == [[0, 1], [1, 2], [0, 2]]
    assert ln.tolist() == [3.0, 4.0, 5.0]
    _, _, c = s.polygon(0)
    assert np.array_equal(bits(c), bits(np.array([f32(6.0) / f32(3.0), f32(4.0) / f32(3.0)], f32)))


def test_polygon_circle_links():
    # polygon.rs:17-82: chords i<->(i+2n/3)%n and i<->(i+n/3)%n, a<b, 2n links, duplicates when 3|n
    s = bo.OracleSolver()
    s.add_polygon_circle(2.0, (50, 50), 6, False)
    ab, ln = s.polygon_links(0)
    assert len(ab) == 12
    assert (ab[:, 0] < ab[:, 1]).all()
    assert ab[0].tolist() == [0, 4] and ab[1].tolist() == [0, 2]


def test_schedule_order_replay_of_disjoint_links_is_identical():
    # any permutation that only reorders vertex-disjoint links leaves the result bit-identical
    rng = np.random.default_rng(1)
    pts = rng.uniform(10, 90, size=(8, 2)).astype(f32)
    a, b = bo.OracleSolver(), bo.OracleSolver()
    for s in (a, b):
        s.add_particles(pts)
        for k, (i, j) in enumerate([(0, 1), (2, 3), (4, 5), (6, 7)]):
            s.add_particle_link(i, j, 3.0 + k)
    b.set_link_order([3, 1, 0, 2])
    for _ in range(5):
        a.update(0.01)
        b.update(0.01)
    assert np.array_equal(bits(a.particles()[0]), bits(b.particles()[0]))
    with pytest.raises(ValueError):
        b.set_link_order([0, 0, 1, 2])


def test_ext_disc_contact_equals_reference_pair_rule_on_a_matching():
    """ext (unpinned) disc contact: for isolated pairs the Jacobi/fixed-point rule must reproduce
    Circle::solve_circle (circle.rs:32-45) bit for bit."""
    rng = np.random.default_rng(5)
    n_pairs = 400
    base = np.stack([10 + 4.0 * (np.arange(n_pairs) % 20), 10 + 4.0 * (np.arange(n_pairs) // 20)], 1)
    ang = rng.uniform(0, 2 * np.pi, n_pairs)
    sep = rng.uniform(0.02, 0.1999, n_pairs)
    a = (base + 0.5 * sep[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1)).astype(f32)
    b = (base - 0.5 * sep[:, None] * np.stack([np.cos(ang), np.sin(ang)], 1)).astype(f32)
    s = bo.OracleSolver()
    s.set_gravity(0, 0)
    s.set_bounds(0, 0, 100, 100)
    pts = np.empty((2 * n_pairs, 2), f32)
    pts[0::2], pts[1::2] = a, b
    s.add_particles(pts)
    s.set_particle_radius(0.1)
    s.set_grid(0, 0, 1 / 0.2, 500, 500)
    s.update(0.01)
    _, prev = s.particles()  # prev = post-contact position (integrate copies pos into prev)
    hits = 0
    for k in range(n_pairs):
        hit, p1, p2 = bo.prim_circle_solve(a[k], b[k], 0.1, 0.1)
        hits += hit
        assert np.array_equal(bits(prev[2 * k]), bits(p1)) and np.array_equal(bits(prev[2 * k + 1]), bits(p2)), k
    assert hits > 300


def test_ref_harness_pipeline():
    """oracle/ref_harness (the cargo harness that pins the oracle against the real crate) end to end with the oracle
    standing in for the crate: scene export, update counts (C1: 240 calls of dt/8 == 30 frames of 8 substeps),
    output format and the bit compare against tests/golden/*.npz"""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "oracle", "ref_harness", "selfcheck.py")], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count(" OK") == 3
