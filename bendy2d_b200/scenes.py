"""Synthetic scenes of BASELINE.json `configs` (exact constructions: SURVEY.md §8.4).

Pure numpy; a `Scene` can be loaded into the GPU `Solver` (load_into) and, by the tests, into the
CPU oracle.  PRNG = splitmix64 -> uniform f32, seed = 0xB2000000 + config number.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

f32 = np.float32


# ------------------------------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed: int):
        self.state = np.uint64(seed)

    def next_u64(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = self.state + idx * np.uint64(0x9E3779B97F4A7C15)
            self.state = z[-1] if n else self.state
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def uniform(self, n: int, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
        u = (self.next_u64(n) >> np.uint64(40)).astype(np.float64) / float(1 << 24)  # 24-bit mantissa
        return (lo + (hi - lo) * u).astype(f32)


def _mag32(dx: np.ndarray, dy: np.ndarray) -> np.ndarray:
    """nalgebra magnitude in f32: sqrt(dx*dx + dy*dy) with every op rounded to f32."""
    dx, dy = dx.astype(f32), dy.astype(f32)
    return np.sqrt((dx * dx).astype(f32) + (dy * dy).astype(f32), dtype=f32)


@dataclass
class Scene:
    name: str
    bounds: tuple  # (x, y, w, h)
    gravity: tuple = (0.0, 98.2)  # solver.rs:36
    dt: float = 1.0 / 120.0       # per update() call
    sub_steps: int = 1
    particle_radius: float = 0.0
    polygon_contact: bool = False
    particles: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), f32))
    links_ab: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.uint32))
    links_len: np.ndarray = field(default_factory=lambda: np.zeros((0,), f32))
    circles_pos: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), f32))
    circles_r: np.ndarray = field(default_factory=lambda: np.zeros((0,), f32))
    polygons: List[np.ndarray] = field(default_factory=list)  # explicit vertex lists (Polygon::new)
    polygons_static: List[bool] = field(default_factory=list)
    timed_substeps: int = 200
    body_of: Optional[np.ndarray] = None  # body (connected component) id per particle, if the generator knows it

    # -- sizes
    @property
    def n_particles(self) -> int:
        return len(self.particles)

    @property
    def n_links(self) -> int:
        return len(self.links_len)

    @property
    def n_polygon_points(self) -> int:
        return int(sum(len(p) for p in self.polygons))

    @property
    def n_points(self) -> int:  # what "particle-substeps" counts
        return self.n_particles + len(self.circles_r) + self.n_polygon_points

    def n_cells(self, grid=None) -> int:
        if not self.particle_radius > 0:
            return 0
        if grid is not None:
            return grid[3] * grid[4]
        h = 2.0 * self.particle_radius
        return int(np.ceil(self.bounds[2] / h) * np.ceil(self.bounds[3] / h))

    def algorithmic_bytes(self, grid=None) -> dict:
        """Compulsory bytes per substep, per kernel class (SURVEY.md §8.4 / BASELINE.md)."""
        n_pts = self.n_points
        n_disc = self.n_particles if self.particle_radius > 0 else 0
        n_poly_links = int(sum(len(p) for p in self.polygons))  # Polygon::new: one perimeter link per vertex
        k4 = self.polygon_contact and len(self.polygons) > 0
        out = {
            "K1_integrate": n_pts * 32,
            "K2_grid": n_disc * 36 + self.n_cells(grid) * 16,  # hash R8 W4 | scatter R4+8 W8+4 | cells 16
            "K2_narrow": n_disc * 20,                          # narrowphase R8+4 W8
            "K3_links": (self.n_links + n_poly_links) * 44,
            "K4_polygon": (self.n_particles * 16 + self.n_polygon_points * 8 + len(self.polygons) * 24) if k4 else 0,
        }
        out["total"] = sum(out.values())
        return out

    # -- loading
    def load_into(self, solver) -> None:
        """Populate a bendy2d_b200.Solver (or anything with the same add_* surface)."""
        from .solver import Bounds

        solver.gravity = np.array(self.gravity, f32)
        solver.bounds = Bounds(np.array(self.bounds[:2], f32), np.array(self.bounds[2:], f32))
        if self.n_particles:
            solver.add_particles(self.particles)
        if self.n_links:
            solver.add_particle_links(self.links_ab, self.links_len)
        if len(self.circles_r):
            solver.add_circles(self.circles_pos, self.circles_r)
        for pts, st in zip(self.polygons, self.polygons_static):
            ab, ln, cen = polygon_new_tables(pts)
            solver.add_polygon_raw(pts, ab, ln, st, cen)
        solver.set_sub_steps(self.sub_steps)
        solver.set_particle_radius(self.particle_radius)
        solver.set_polygon_contact(self.polygon_contact)


def polygon_new_tables(pts: np.ndarray):
    """Polygon::new (polygon.rs:84-123) for one explicit vertex list: links (a<b), lengths, centre."""
    pts = np.asarray(pts, f32)
    n = len(pts)
    cx, cy = f32(0), f32(0)
    for p in pts:  # sequential f32 sum, then / n
        cx, cy = f32(cx + p[0]), f32(cy + p[1])
    cen = np.array([cx / f32(n), cy / f32(n)], f32)
    a = np.arange(n)
    b = (a + 1) % n
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    d = pts[lo] - pts[hi]
    return np.stack([lo, hi], 1).astype(np.uint32), _mag32(d[:, 0], d[:, 1]), cen


# ------------------------------------------------------------------------------------------------
def lattice_body(cols: int, rows: int, spacing: float, origin, both_diagonals: bool):
    """Row-major lattice; links emitted per particle: right, down, down-right[, down-left]; a<b."""
    ox, oy = origin
    c, r = np.meshgrid(np.arange(cols), np.arange(rows))
    x = (f32(ox) + c.astype(f32) * f32(spacing)).astype(f32)
    y = (f32(oy) + r.astype(f32) * f32(spacing)).astype(f32)
    pos = np.stack([x.ravel(), y.ravel()], 1).astype(f32)
    idx = (r * cols + c)
    cand = []  # (order key = source particle, slot), a, b
    def add(mask, da_r, da_c, slot):
        src = idx[mask]
        dst = (r[mask] + da_r) * cols + (c[mask] + da_c)
        cand.append((src * 4 + slot, np.minimum(src, dst), np.maximum(src, dst)))
    add(c < cols - 1, 0, 1, 0)
    add(r < rows - 1, 1, 0, 1)
    add((c < cols - 1) & (r < rows - 1), 1, 1, 2)
    if both_diagonals:
        add((c > 0) & (r < rows - 1), 1, -1, 3)
    key = np.concatenate([k for k, _, _ in cand])
    a = np.concatenate([x_ for _, x_, _ in cand])
    b = np.concatenate([x_ for _, _, x_ in cand])
    order = np.argsort(key, kind="stable")
    return pos, np.stack([a[order], b[order]], 1).astype(np.uint32)


def _replicate_bodies(pos0, ab0, offsets):
    """Copies of one body translated by `offsets` (f32 adds), links re-based."""
    nb, npb = len(offsets), len(pos0)
    off = np.asarray(offsets, f32)
    pos = (pos0[None, :, :] + off[:, None, :]).astype(f32).reshape(-1, 2)
    ab = (ab0[None, :, :].astype(np.int64) + (np.arange(nb, dtype=np.int64) * npb)[:, None, None]).reshape(-1, 2)
    d = pos[ab[:, 0]] - pos[ab[:, 1]]
    return pos, ab.astype(np.uint32), _mag32(d[:, 0], d[:, 1])


def regular_polygon(cx, cy, radius, k, theta0):
    ang = theta0 + 2.0 * np.pi * np.arange(k) / k
    return np.stack([cx + radius * np.cos(ang), cy + radius * np.sin(ang)], 1).astype(f32)


# ------------------------------------------------------------------------------------------------
def c1_softbody_blob() -> Scene:
    """C1: 20x20 lattice, spacing 1, both diagonals: 400 particles, 1482 links, one Circle r=5.
    Runs literally on the reference semantics (all extensions off); dt=1/60 with 8 substeps."""
    pos, ab = lattice_body(20, 20, 1.0, (40.0, 10.0), True)
    d = pos[ab[:, 0]] - pos[ab[:, 1]]
    return Scene("C1 softbody blob 20x20", (0.0, 0.0, 100.0, 100.0), dt=1.0 / 60.0, sub_steps=8,
                 particles=pos, links_ab=ab, links_len=_mag32(d[:, 0], d[:, 1]),
                 circles_pos=np.array([[50.0, 70.0]], f32), circles_r=np.array([5.0], f32), timed_substeps=8000)


def c2_free_particles(n_cols: int = 400, n_rows: int = 250) -> Scene:
    """C2: 100k discs r=0.1 on a jittered grid (pitch 0.25, jitter +-0.02) at the top of (0,0)+(128,128)."""
    rng = SplitMix64(0xB2000000 + 2)
    c, r = np.meshgrid(np.arange(n_cols), np.arange(n_rows))
    n = n_cols * n_rows
    jx, jy = rng.uniform(n, -0.02, 0.02), rng.uniform(n, -0.02, 0.02)
    x = (f32(14.0) + c.ravel().astype(f32) * f32(0.25) + jx).astype(f32)
    y = (f32(2.0) + r.ravel().astype(f32) * f32(0.25) + jy).astype(f32)
    return Scene(f"C2 {n} free particles", (0.0, 0.0, 128.0, 128.0), particle_radius=0.1,
                 particles=np.stack([x, y], 1), timed_substeps=1000)


def softbody_field(bodies_x: int, bodies_y: int, bounds, origin, n_circles: int, n_polygons: int, seed: int,
                   name: str, timed: int) -> Scene:
    pos0, ab0 = lattice_body(20, 25, 0.25, (0.0, 0.0), False)  # 500 particles, 1411 links
    bx, by = np.meshgrid(np.arange(bodies_x), np.arange(bodies_y))
    offs = np.stack([origin[0] + 8.0 * bx.ravel(), origin[1] + 9.0 * by.ravel()], 1)
    pos, ab, ln = _replicate_bodies(pos0, ab0, offs)
    rng = SplitMix64(seed)
    y_bodies_end = origin[1] + 9.0 * bodies_y
    circles_pos = np.zeros((0, 2), f32)
    circles_r = np.zeros((0,), f32)
    polys, statics = [], []
    if n_circles:
        per_row = 20
        rows = (n_circles + per_row - 1) // per_row
        k = np.arange(n_circles)
        cx = 16.0 + (bounds[2] - 32.0) * ((k % per_row) + 0.5) / per_row + rng.uniform(n_circles, -1.0, 1.0)
        cy = y_bodies_end + 4.0 + 7.0 * (k // per_row) + rng.uniform(n_circles, -0.25, 0.25)
        circles_pos = np.stack([cx, cy], 1).astype(f32)
        circles_r = rng.uniform(n_circles, 1.0, 3.0)
        y_poly0 = y_bodies_end + 4.0 + 7.0 * rows + 3.0
    else:
        y_poly0 = y_bodies_end + 6.0
    if n_polygons:
        per_row = 50
        k = np.arange(n_polygons)
        px = 16.0 + (bounds[2] - 32.0) * ((k % per_row) + 0.5) / per_row
        py = y_poly0 + 7.0 * (k // per_row)
        th = rng.uniform(n_polygons, 0.0, 1.0).astype(np.float64)
        for i in range(n_polygons):
            polys.append(regular_polygon(px[i], py[i], 3.0, 6, th[i]))
            statics.append(True)
    return Scene(name, bounds, particle_radius=0.1, polygon_contact=n_polygons > 0, particles=pos, links_ab=ab,
                 links_len=ln, circles_pos=circles_pos, circles_r=circles_r, polygons=polys, polygons_static=statics,
                 timed_substeps=timed, body_of=np.repeat(np.arange(len(offs), dtype=np.int64), len(pos0)))


def c3_softbody_field(bodies_x: int = 50, bodies_y: int = 40, n_circles: int = 200, n_polygons: int = 500) -> Scene:
    """C3: 2,000 bodies x (20x25 lattice, d=0.25) = 1,000,000 particles, 2,822,000 links (h+v+one
    diagonal), 200 Circles r in U(1,3), 500 static 6-gons, bounds (0,0)+(512,512), r_p = 0.1."""
    return softbody_field(bodies_x, bodies_y, (0.0, 0.0, 512.0, 512.0), (56.0, 8.0), n_circles, n_polygons,
                          0xB2000000 + 3, f"C3 softbody field {bodies_x * bodies_y} bodies", 200)


def c4_polygon_heavy(grid: int = 100, per_band: int = 2000) -> Scene:
    """C4: 200k free discs r=0.1 raining over 10,000 static convex polygons (4-8 vertices,
    circumradius U(1,2.5), jittered grid pitch 6) in bounds (0,0)+(640,640)."""
    rng = SplitMix64(0xB2000000 + 4)
    n_poly = grid * grid
    gx, gy = np.meshgrid(np.arange(grid), np.arange(grid))
    cx = 23.0 + 6.0 * gx.ravel() + rng.uniform(n_poly, -0.4, 0.4)
    cy = 23.0 + 6.0 * gy.ravel() + rng.uniform(n_poly, -0.4, 0.4)
    rad = rng.uniform(n_poly, 1.0, 2.5)
    kk = 4 + (rng.next_u64(n_poly) % np.uint64(5)).astype(np.int64)
    th = rng.uniform(n_poly, 0.0, 6.2831853).astype(np.float64)
    polys = [regular_polygon(cx[i], cy[i], float(rad[i]), int(kk[i]), th[i]) for i in range(n_poly)]
    n = grid * per_band
    band, col = np.meshgrid(np.arange(grid), np.arange(per_band), indexing="ij")
    span = 6.0 * grid
    x = 20.0 + span * (col.ravel() + 0.5) / per_band + rng.uniform(n, -0.02, 0.02)
    y = 20.0 + 6.0 * band.ravel() + rng.uniform(n, -0.3, 0.3)
    size = 40.0 + span
    return Scene(f"C4 polygon-heavy {n} particles / {n_poly} polygons", (0.0, 0.0, size, size), particle_radius=0.1,
                 polygon_contact=True, particles=np.stack([x, y], 1).astype(f32), polygons=polys,
                 polygons_static=[True] * n_poly, timed_substeps=200)


def c5_softbody_field_16m(bodies_x: int = 200, bodies_y: int = 160) -> Scene:
    """C5: 32,000 bodies x 500 = 16,000,000 particles, 45,152,000 links, bounds (0,0)+(2048,2048)."""
    return softbody_field(bodies_x, bodies_y, (0.0, 0.0, 2048.0, 2048.0), (224.0, 8.0), 0, 0, 0xB2000000 + 5,
                          f"C5 softbody field {bodies_x * bodies_y} bodies", 100)


