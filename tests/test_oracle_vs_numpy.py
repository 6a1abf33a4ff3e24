"""Randomised cross-check of the C++ oracle's primitives against an independent numpy-float32 restatement
(tests/np_restatement.py).  The reference has no tests and cannot be built here ("parity unpinned"), so
besides the hand-derived KATs the strongest available pin is two restatements, written separately from the
Rust source in different languages, agreeing bit for bit on thousands of random inputs (SURVEY.md section 4:
randomised single-primitive checks)."""
import numpy as np

import np_restatement as R
from oracle import bo

from helpers import bits

f32 = np.float32
N = 4000


def same(a, b):
    a, b = np.asarray(a, f32).ravel(), np.asarray(b, f32).ravel()
    return np.array_equal(bits(a), bits(b)) or (np.isnan(a) == np.isnan(b)).all() and np.array_equal(
        bits(a)[~np.isnan(a)], bits(b)[~np.isnan(b)])


def rnd(rng, scale=100.0):
    """Mostly generic points, sometimes axis-aligned / coincident / huge / tiny ones."""
    p = rng.uniform(-scale, scale, 2).astype(f32)
    k = rng.integers(0, 12)
    if k == 0:
        p[0] = 0.0
    elif k == 1:
        p = np.round(p)
    elif k == 2:
        p *= f32(1e-20)
    elif k == 3:
        p *= f32(1e18)
    return p


def test_particle_link_solve():
    rng = np.random.default_rng(1)
    for i in range(N):
        a, b = rnd(rng), rnd(rng)
        if i % 50 == 0:
            b = a.copy()  # K-nan: coincident ends
        if i % 7 == 0:
            b[1] = a[1]  # axis-aligned (the zero-numerator shortcut of the device code)
        L = f32(rng.uniform(0, 50))
        with np.errstate(all="ignore"):
            ra, rb = R.particle_link_solve(tuple(a), tuple(b), L)
        oa, ob = bo.prim_link_solve(a, b, float(L))
        assert same(ra, oa) and same(rb, ob), (i, a, b, L)


def test_circle_link_solve():
    rng = np.random.default_rng(2)
    for i in range(N):
        a, b = rnd(rng), rnd(rng)
        ra_, rb_, L = f32(rng.uniform(0.1, 5)), f32(rng.uniform(0.1, 5)), f32(rng.uniform(0, 50))
        with np.errstate(all="ignore"):
            xa, xb = R.circle_link_solve(tuple(a), tuple(b), ra_, rb_, L)
        oa, ob = bo.prim_circle_link_solve(a, b, float(ra_), float(rb_), float(L))
        assert same(xa, oa) and same(xb, ob), (i, a, b)


def test_circle_solve():
    rng = np.random.default_rng(3)
    hits = 0
    for i in range(N):
        a = rnd(rng, 20.0)
        b = (a + rng.uniform(-3, 3, 2)).astype(f32) if i % 2 else rnd(rng, 20.0)
        if i % 100 == 0:
            b = a.copy()
        r1, r2 = f32(rng.uniform(0.1, 3)), f32(rng.uniform(0.1, 3))
        with np.errstate(all="ignore"):
            h, xa, xb = R.circle_solve(tuple(a), tuple(b), r1, r2)
        oh, oa, ob = bo.prim_circle_solve(a, b, float(r1), float(r2))
        assert h == oh and same(xa, oa) and same(xb, ob), (i, a, b, r1, r2)
        hits += h
    assert hits > N // 10
    # strict '<' (circle.rs:36): touching discs are left alone by both
    h, _, _ = R.circle_solve((f32(0), f32(0)), (f32(2), f32(0)), f32(1), f32(1))
    oh, _, _ = bo.prim_circle_solve(np.array([0, 0], f32), np.array([2, 0], f32), 1.0, 1.0)
    assert not h and not oh


def test_particle_update_and_bounds():
    rng = np.random.default_rng(4)
    bounds = (f32(1.5), f32(-2.25), f32(100.1), f32(63.7))
    for i in range(N):
        pos = rnd(rng, 120.0)
        prev = (pos + rng.uniform(-1, 1, 2)).astype(f32)
        acc = rng.uniform(-200, 200, 2).astype(f32)
        dt = f32(rng.uniform(1e-4, 0.05))
        xp, xq, xa = R.particle_update(tuple(pos), tuple(prev), tuple(acc), dt)
        op, oq, oa = bo.prim_particle_update(pos, prev, acc, float(dt))
        assert same(xp, op) and same(xq, oq) and same(xa, oa), (i, pos, prev, acc, dt)
        xp, xq = R.particle_bounds(tuple(pos), tuple(prev), bounds)
        op, oq = bo.prim_particle_bounds(pos, prev, [float(x) for x in bounds])
        assert same(xp, op) and same(xq, oq), (i, pos, prev)
        r = f32(rng.uniform(0.1, 4))
        xp, xq = R.circle_bounds(tuple(pos), tuple(prev), r, bounds)
        op, oq = bo.prim_circle_bounds(pos, prev, float(r), [float(x) for x in bounds])
        assert same(xp, op) and same(xq, oq), (i, pos, prev, r)


def test_line_intersection():
    rng = np.random.default_rng(5)
    hits = 0
    for i in range(N):
        p = [rnd(rng, 10.0) for _ in range(4)]
        if i % 20 == 0:
            p[3] = (p[2] + (p[1] - p[0])).astype(f32)  # K-par: parallel segments
        if i % 33 == 0:
            p[2] = p[0].copy()  # shared end point
        x = R.line_intersection(*[tuple(q) for q in p])
        o = bo.prim_line_intersection(*p)
        assert (x is None) == (o is None), (i, p)
        if x is not None:
            assert same(x, o), (i, p)
            hits += 1
    assert hits > N // 10


def convex_polygon(rng, centre, radius, k):
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    pts = np.stack([centre[0] + radius * np.cos(ang), centre[1] + radius * np.sin(ang)], 1).astype(f32)
    return pts


def test_polygon_contact_chain():
    """resolve_line_intersection and the whole solve_polygon_single nest (stale edge copies, last writer
    wins, polygon.rs:147-216) on random overlapping convex polygons."""
    rng = np.random.default_rng(6)
    touched = 0
    for i in range(600):
        ca = rng.uniform(-5, 5, 2)
        cb = ca + rng.uniform(-3, 3, 2)
        A = convex_polygon(rng, ca, rng.uniform(1, 3), rng.integers(3, 8))
        B = convex_polygon(rng, cb, rng.uniform(1, 3), rng.integers(3, 8))
        cA, cB = R.calc_center([tuple(p) for p in A]), R.calc_center([tuple(p) for p in B])
        # single edge x point first
        q = B[rng.integers(0, len(B))]
        x = R.resolve_line_intersection(cA, tuple(A[0]), tuple(A[1]), tuple(q), cB)
        o = bo.prim_resolve_line_intersection(A[0], A[1], q, np.array(cB, f32), np.array(cA, f32))
        assert (x is None) == (o is None), i
        if x is not None:
            assert same(np.array(x), o), i
        xs, xo = R.solve_polygon_single(A, cA, B, cB)
        os_, oo = bo.prim_solve_polygon_single(A, np.array(cA, f32), B, np.array(cB, f32))
        assert same(np.array(xs), os_) and same(np.array(xo), oo), i
        touched += not np.array_equal(bits(os_), bits(A))
    assert touched > 100


def test_whole_update_sequence_on_a_small_mixed_scene():
    """Phase order, loop order and the static-polygon rules of Solver::update (solver.rs:106-188): the numpy
    restatement and the C++ oracle, both in reference mode, stay bit-identical over 60 updates of a scene
    with links, circle piles, a circle link, overlapping dynamic polygons, a static polygon and wall hits."""
    from bendy2d_b200 import scenes

    rng = np.random.default_rng(11)
    bounds = (0.0, 0.0, 24.0, 18.0)
    n = R.NpSolver(bounds=bounds, sub_steps=2)
    o = bo.OracleSolver()
    o.set_bounds(*bounds)
    o.set_sub_steps(2)
    pos, ab = scenes.lattice_body(4, 4, 1.0, (3.0, 2.0), True)
    ln = np.sqrt(((pos[ab[:, 0]].astype(np.float64) - pos[ab[:, 1]]) ** 2).sum(1)).astype(f32)
    pos = (pos + rng.uniform(-0.1, 0.1, pos.shape)).astype(f32)
    for p in pos:
        n.add_particle(p)
        o.add_particle(float(p[0]), float(p[1]))
    for (a, b), l in zip(ab, ln):
        n.particle_links.append((int(a), int(b), f32(l)))
        o.add_particle_link(int(a), int(b), float(l))
    cpos = rng.uniform(8, 16, (6, 2)).astype(f32)
    crad = rng.uniform(0.8, 2.0, 6).astype(f32)
    for p, r in zip(cpos, crad):
        n.add_circle(p, r)
        o.add_circle(p, float(r))
    n.circle_links.append((1, 4, f32(3.0)))
    o.add_circle_link(1, 4, 3.0)
    polys = [(scenes.regular_polygon(18.0, 6.0, 2.0, 5, 0.3), False),
             (scenes.regular_polygon(19.5, 7.0, 2.0, 4, 0.1), False),
             (scenes.regular_polygon(19.0, 15.0, 2.5, 6, 0.0), True)]
    for pts, st in polys:
        pts = np.asarray(pts, f32)
        pab, pln, cen = scenes.polygon_new_tables(pts)
        n.add_polygon(pts, [(a, b, l) for (a, b), l in zip(pab, pln)], st, cen)
        o.add_polygon(pts, pab, pln, st, cen)
    ymax = f32(0)
    for k in range(60):
        n.update(1 / 60)
        o.update(1 / 60)
        ymax = max(ymax, max(p[0][1] for p in n.particles))
        op, oq = o.particles()
        assert same([p[0] for p in n.particles], op) and same([p[1] for p in n.particles], oq), k
        cp, cq, _ = o.circles()
        assert same([c[0] for c in n.circles], cp) and same([c[1] for c in n.circles], cq), k
        for g in range(3):
            pp, pq, pc = o.polygon(g)
            assert same([p[0] for p in n.polygons[g]["points"]], pp), (k, g)
            assert same([p[1] for p in n.polygons[g]["points"]], pq), (k, g)
            assert same(n.polygons[g]["center"], pc), (k, g)
    # the scene really exercised the walls and the polygon contact
    assert ymax >= f32(17.5)
    assert not same([p[0] for p in n.polygons[0]["points"]], np.asarray(polys[0][0], f32))
