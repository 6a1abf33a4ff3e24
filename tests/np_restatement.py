"""A second, independent restatement of the reference's primitives in numpy float32 scalars.

Test infrastructure (like oracle/): written straight from the Rust sources, sharing no code with the C++
oracle, so that the randomised cross-checks in test_oracle_vs_numpy.py compare two implementations that can
only agree if both follow the reference's operator order.  Every numpy float32 operation rounds once to f32
(no fused multiply-add, no extended precision), which is what rustc emits for these lines.

nalgebra 0.32 semantics assumed: dot = x*x' + y*y' (two rounded products, one rounded add),
magnitude = sqrt(dot(v, v)), normalize = component-wise division by the magnitude, vector*scalar and
scalar*vector component-wise, expressions left-associative as written.
"""
import numpy as np

F = np.float32
ZERO, ONE, HALF = F(0.0), F(1.0), F(0.5)


def v(x, y):
    return (F(x), F(y))


def add(a, b):
    return (F(a[0] + b[0]), F(a[1] + b[1]))


def sub(a, b):
    return (F(a[0] - b[0]), F(a[1] - b[1]))


def scale(a, s):
    return (F(a[0] * s), F(a[1] * s))


def div(a, s):
    return (F(a[0] / s), F(a[1] / s))


def dot(a, b):
    return F(F(a[0] * b[0]) + F(a[1] * b[1]))


def magnitude(a):
    return F(np.sqrt(dot(a, a)))


def normalize(a):
    return div(a, magnitude(a))


def particle_link_solve(a, b, target):  # link.rs:18-27
    dist_vec = sub(a, b)
    dist = magnitude(dist_vec)
    normal = normalize(dist_vec)
    corr = scale(scale(normal, F(dist - target)), HALF)
    return sub(a, corr), add(b, corr)


def circle_link_solve(a, b, ra, rb, target):  # link.rs:36-48
    dist_vec = sub(a, b)
    dist = magnitude(dist_vec)
    normal = normalize(dist_vec)
    a2, b2 = F(ra * ra), F(rb * rb)
    s = F(ONE / F(a2 + b2))
    base = scale(scale(normal, F(dist - target)), s)
    return sub(a, scale(base, b2)), add(b, scale(base, a2))


def circle_solve(p1, p2, r1, r2):  # circle.rs:32-45
    dist = sub(p1, p2)
    dist_sqr = dot(dist, dist)
    radius_sum = F(r1 + r2)
    if not dist_sqr < F(radius_sum * radius_sum):
        return False, p1, p2
    normal = normalize(dist)
    overlap = F(radius_sum - F(np.sqrt(dist_sqr)))
    s1, s2 = F(r1 * r1), F(r2 * r2)
    sc = F(ONE / F(s1 + s2))
    base = scale(scale(normal, sc), overlap)
    return True, add(p1, scale(base, s2)), sub(p2, scale(base, s1))


def particle_update(pos, prev, acc, dt):  # particle.rs:20-25
    vel = sub(pos, prev)
    new_prev = pos
    new_pos = add(add(pos, vel), scale(scale(acc, dt), dt))
    return new_pos, new_prev, v(0, 0)


def _axis_bounds(p, q, lo, size, inset):
    # particle.rs:28-45 / circle.rs:12-29: `lo + size` (and `lo + size - r`) are recomputed at every use
    if p < F(lo + inset):
        vel = F(q - p)
        return F(lo + inset), F(F(lo + inset) - vel)
    if p > F(F(lo + size) - inset):
        vel = F(q - p)
        return F(F(lo + size) - inset), F(F(F(lo + size) - inset) - vel)
    return p, q


def particle_bounds(pos, prev, b):  # particle.rs:27-46; b = (x, y, w, h)
    # no radius: the reference compares against bounds.pos.x and bounds.pos.x + bounds.size.x directly
    def axis(p, q, lo, size):
        if p < lo:
            vel = F(q - p)
            return lo, F(lo - vel)
        if p > F(lo + size):
            vel = F(q - p)
            return F(lo + size), F(F(lo + size) - vel)
        return p, q

    px, qx = axis(pos[0], prev[0], F(b[0]), F(b[2]))
    py, qy = axis(pos[1], prev[1], F(b[1]), F(b[3]))
    return (px, py), (qx, qy)


def circle_bounds(pos, prev, r, b):  # circle.rs:11-30
    px, qx = _axis_bounds(pos[0], prev[0], F(b[0]), F(b[2]), F(r))
    py, qy = _axis_bounds(pos[1], prev[1], F(b[1]), F(b[3]), F(r))
    return (px, py), (qx, qy)


def line_intersection(p1, p2, p3, p4):  # common.rs:4-26
    s1x, s1y = F(p2[0] - p1[0]), F(p2[1] - p1[1])
    s2x, s2y = F(p4[0] - p3[0]), F(p4[1] - p3[1])
    with np.errstate(all="ignore"):
        s = F(F(F(-s1y * F(p1[0] - p3[0])) + F(s1x * F(p1[1] - p3[1]))) / F(F(-s2x * s1y) + F(s1x * s2y)))
        t = F(F(F(s2x * F(p1[1] - p3[1])) - F(s2y * F(p1[0] - p3[0]))) / F(F(-s2x * s1y) + F(s1x * s2y)))
    if s >= ZERO and s <= ONE and t >= ZERO and t <= ONE:
        return (F(p1[0] + F(t * s1x)), F(p1[1] + F(t * s1y)))
    return None


def resolve_line_intersection(self_center, pa, pb, q, other_center):  # polygon.rs:164-216
    hit = line_intersection(pa, pb, q, other_center)
    if hit is None:
        return None
    with np.errstate(all="ignore"):
        normal_line = normalize(sub(pb, pa))
        center_proj = scale(normal_line, F(dot(normal_line, sub(self_center, hit)) / dot(normal_line, normal_line)))
        normal_in = normalize(sub(self_center, add(hit, center_proj)))
        dist_to_a, dist_to_b = magnitude(sub(hit, pa)), magnitude(sub(hit, pb))
        dist_a_to_b = F(dist_to_a + dist_to_b)
        influence_a, influence_b = F(dist_to_b / dist_a_to_b), F(dist_to_a / dist_a_to_b)
        diff = sub(hit, q)
        on_normal = scale(normal_in, F(dot(normal_in, diff) / dot(normal_in, normal_in)))
        displace_line = scale(div(on_normal, F(3.0)), F(2.0))
        new_a = sub(pa, scale(displace_line, influence_a))
        new_b = sub(pb, scale(displace_line, influence_b))
        new_q = line_intersection(pa, pb, q, sub(q, scale(normal_in, F(10000.0))))
    if new_q is None:
        return None
    return new_a, new_b, new_q


def solve_polygon_single(self_pts, self_center, other_pts, other_center):  # polygon.rs:147-162
    self_pts, other_pts = [tuple(p) for p in self_pts], [tuple(p) for p in other_pts]
    n = len(self_pts)
    for i in range(n):
        pa, b_id = self_pts[i], (i + 1) % n  # per-edge COPIES: stale across the inner loop
        pb = self_pts[b_id]
        for k in range(len(other_pts)):
            r = resolve_line_intersection(self_center, pa, pb, other_pts[k], other_center)
            if r is not None:
                self_pts[i], self_pts[b_id], other_pts[k] = r
    return self_pts, other_pts


def calc_center(pts):  # polygon.rs:231-237: sequential sum, then one division
    c = v(0, 0)
    for p in pts:
        c = add(c, p)
    return div(c, F(len(pts)))


class NpSolver:
    """Solver::update and what it calls (solver.rs:106-188, polygon.rs:125-140,218-237), reference semantics
    only (no extensions).  Lists of mutable records; quadratic loops exactly as written."""

    def __init__(self, gravity=(0.0, 98.2), bounds=(0.0, 0.0, 100.0, 100.0), sub_steps=1):
        self.gravity, self.bounds, self.sub_steps = v(*gravity), tuple(F(x) for x in bounds), sub_steps
        self.particles, self.circles, self.polygons = [], [], []  # [pos, prev, acc] (+ radius for circles)
        self.particle_links, self.circle_links = [], []

    def add_particle(self, pos):
        self.particles.append([v(*pos), v(*pos), v(0, 0)])

    def add_circle(self, pos, radius, prev=None, acc=(0.0, 0.0)):
        self.circles.append([v(*pos), v(*(pos if prev is None else prev)), v(*acc), F(radius)])

    def add_polygon(self, pts, links, is_static, center):
        self.polygons.append({"points": [[v(*p), v(*p), v(0, 0)] for p in pts],
                              "links": [(int(a), int(b), F(l)) for a, b, l in links],
                              "is_static": bool(is_static), "center": v(*center)})

    # -- solver.rs:106-116
    def update(self, dt):
        mult = F(ONE / F(self.sub_steps))
        delta = F(F(dt) * mult)
        for _ in range(self.sub_steps):
            self.apply_gravity()
            self.apply_links()
            self.solve_dynamic_collisions()
            self.solve_boundary_collisions()
            self.update_positions(delta)

    def apply_gravity(self):  # solver.rs:130-141 (static polygons included, polygon.rs:225-229)
        for p in self.particles + self.circles:
            p[2] = add(p[2], self.gravity)
        for g in self.polygons:
            for p in g["points"]:
                p[2] = add(p[2], self.gravity)

    def apply_links(self):  # solver.rs:143-153
        for a, b, L in self.particle_links:
            self.particles[a][0], self.particles[b][0] = particle_link_solve(self.particles[a][0],
                                                                               self.particles[b][0], L)
        for a, b, L in self.circle_links:
            ca, cb = self.circles[a], self.circles[b]
            ca[0], cb[0] = circle_link_solve(ca[0], cb[0], ca[3], cb[3], L)
        for g in self.polygons:  # polygon.rs:218-223
            g["center"] = calc_center([p[0] for p in g["points"]])
            for a, b, L in g["links"]:
                g["points"][a][0], g["points"][b][0] = particle_link_solve(g["points"][a][0], g["points"][b][0], L)

    def solve_dynamic_collisions(self):  # solver.rs:167-188
        n = len(self.circles)
        for i in range(n):
            for j in range(i + 1, n):
                ci, cj = self.circles[i], self.circles[j]
                with np.errstate(all="ignore"):
                    _, ci[0], cj[0] = circle_solve(ci[0], cj[0], ci[3], cj[3])
        n = len(self.polygons)
        for i in range(n):
            for j in range(i + 1, n):
                gi, gj = self.polygons[i], self.polygons[j]
                for s_, o_ in ((gi, gj), (gj, gi)):  # polygon.rs:142-145
                    sp, op = solve_polygon_single([p[0] for p in s_["points"]], s_["center"],
                                                  [p[0] for p in o_["points"]], o_["center"])
                    for p, new in zip(s_["points"], sp):
                        p[0] = new
                    for p, new in zip(o_["points"], op):
                        p[0] = new

    def solve_boundary_collisions(self):  # solver.rs:155-165
        for p in self.particles:
            p[0], p[1] = particle_bounds(p[0], p[1], self.bounds)
        for c in self.circles:
            c[0], c[1] = circle_bounds(c[0], c[1], c[3], self.bounds)
        for g in self.polygons:
            for p in g["points"]:
                p[0], p[1] = particle_bounds(p[0], p[1], self.bounds)

    def update_positions(self, dt):  # solver.rs:118-128, polygon.rs:125-134
        for p in self.particles + self.circles:
            p[0], p[1], p[2] = particle_update(p[0], p[1], p[2], dt)
        for g in self.polygons:
            if g["is_static"]:
                continue
            g["center"] = calc_center([p[0] for p in g["points"]])
            for p in g["points"]:
                p[0], p[1], p[2] = particle_update(p[0], p[1], p[2], dt)
