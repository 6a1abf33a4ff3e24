"""Host-side mirror of bendy2d's public API over the C ABI of libbendy2d_b200.so.

Same names, argument meaning and error behaviour as the reference's `Solver`, `Particle`, `Link`,
`ParticleLink`, `CircleLink`, `Circle`, `Polygon`, `Bounds` (reference src/solver.rs:13-116,
src/particle.rs:5-18, src/link.rs:5-15,30-33, src/circle.rs:5-8, src/polygon.rs:8-123).  The Rust
façade (rust/bendy2d) binds the same C symbols; this module is the Python spelling of it, used by
the parity tests and bench.py.  All state lives on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import f32p, u32p


class BendyError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_lib.ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class LinkPanic(BendyError):
    """Raised where the reference panics: link with a >= b or b out of range (link.rs:19-21)."""


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _u(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _fp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(f32p)


def _up(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(u32p)


# ------------------------------------------------------------------------------------------------
# plain data types (reference field names)
@dataclass
class Bounds:  # solver.rs:13-17
    pos: np.ndarray
    size: np.ndarray


@dataclass
class Particle:  # particle.rs:5-18
    pos: np.ndarray
    prev_pos: np.ndarray = None
    acc: np.ndarray = field(default_factory=lambda: np.zeros(2, np.float32))

    def __post_init__(self):
        self.pos = np.asarray(self.pos, np.float32)
        self.prev_pos = self.pos.copy() if self.prev_pos is None else np.asarray(self.prev_pos, np.float32)

    @staticmethod
    def new(pos) -> "Particle":
        return Particle(pos)

    def add_force(self, fx: float, fy: float):  # particle.rs:48-50
        self.acc = (self.acc + np.array([fx, fy], np.float32)).astype(np.float32)

    def add_force_v2(self, force):  # particle.rs:52-54
        self.acc = (self.acc + np.asarray(force, np.float32)).astype(np.float32)

    def add_force_towards(self, point, force: float):  # particle.rs:56-59: acc += normalize(point - pos) * force
        d = (np.asarray(point, np.float32) - self.pos).astype(np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            n = (d / _magnitude(d)).astype(np.float32)  # coincident points give NaN, like the reference
            self.acc = (self.acc + n * np.float32(force)).astype(np.float32)


@dataclass
class Link:  # link.rs:5-10
    particle_a: int
    particle_b: int
    target_distance: float


@dataclass
class ParticleLink:  # link.rs:12-15
    link: Link


@dataclass
class CircleLink:  # link.rs:30-33
    link: Link


@dataclass
class Circle:  # circle.rs:5-8
    point: Particle
    radius: float


def _magnitude(d: np.ndarray) -> np.float32:
    d = d.astype(np.float32)
    return np.sqrt(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1]), dtype=np.float32)


@dataclass
class Polygon:  # polygon.rs:8-14
    particles: List[Particle]
    particle_links: List[ParticleLink]
    is_static: bool
    center: np.ndarray
    scale: float = 1.0

    @staticmethod
    def new(points: Sequence, is_static: bool) -> "Polygon":
        """Polygon::new (polygon.rs:84-123): perimeter links, rest length = initial distance."""
        pts = [Particle(np.asarray(p, np.float32)) for p in points]
        center = np.zeros(2, np.float32)
        for p in pts:
            center = (center + p.pos).astype(np.float32)
        center = (center / np.float32(len(pts))).astype(np.float32)
        links = []
        n = len(pts)
        for i in range(n):
            a, b = i, (i + 1) % n
            if a > b:
                a, b = b, a
            links.append(ParticleLink(Link(a, b, float(_magnitude(pts[a].pos - pts[b].pos)))))
        return Polygon(pts, links, is_static, center, 1.0)

    @staticmethod
    def circle(radius: float, pos, point_count: int, is_static: bool) -> "Polygon":
        """Polygon::circle (polygon.rs:17-82): f32 angle accumulation; chords i<->i+2n/3, i<->i+n/3."""
        pos = np.asarray(pos, np.float32)
        pts = []
        center = np.zeros(2, np.float32)
        angle = np.float32(0.0)
        step = np.float32(np.float32(2.0) * np.float32(math.pi)) / np.float32(point_count)
        for _ in range(point_count):
            # f32::cos / f32::sin (polygon.rs:23-24): single-precision libm, not the double result rounded
            x = np.float32(radius) * np.cos(angle, dtype=np.float32)
            y = np.float32(radius) * np.sin(angle, dtype=np.float32)
            p = Particle((pos + np.array([x, y], np.float32)).astype(np.float32))
            pts.append(p)
            center = (center + p.pos).astype(np.float32)
            angle = np.float32(angle + step)
        center = (center / np.float32(point_count)).astype(np.float32)
        links = []
        n = point_count
        for i in range(n):
            for b in ((i + 2 * n // 3) % n, (i + n // 3) % n):
                a = i
                if a > b:
                    a, b = b, a
                links.append(ParticleLink(Link(a, b, float(_magnitude(pts[a].pos - pts[b].pos)))))
        return Polygon(pts, links, is_static, center, 1.0)


class ParticleArray:
    """What get_particles()/get_circles() return: SoA numpy views plus list-like access."""

    def __init__(self, pos: np.ndarray, prev: np.ndarray, radius: Optional[np.ndarray] = None):
        self.pos, self.prev_pos, self.radius = pos, prev, radius

    def __len__(self):
        return len(self.pos)

    def __getitem__(self, i):
        p = Particle(self.pos[i].copy(), self.prev_pos[i].copy())
        return p if self.radius is None else Circle(p, float(self.radius[i]))


# ------------------------------------------------------------------------------------------------
class Solver:
    """Drop-in for bendy2d::solver::Solver (solver.rs:20-116) running on one B200."""

    def __init__(self, device: int = -1, _handle=None):
        self._L = _lib.lib()
        self._h = _handle if _handle is not None else self._L.bendy_create(device)
        if not self._h:
            msg = (self._L.bendy_last_error(None) or b"").decode()
            raise BendyError(-5, msg or "bendy_create failed")
        # Solver::new defaults, solver.rs:34-50
        self.gravity = np.array([0.0, 98.2], np.float32)
        self.bounds = Bounds(np.array([0.0, 0.0], np.float32), np.array([100.0, 100.0], np.float32))
        self.bounds_active = True  # never read by the reference (solver.rs:155-165)

    @staticmethod
    def new(device: int = -1) -> "Solver":
        return Solver(device)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.bendy_destroy(h)

    def _ck(self, rc: int):
        if rc != _lib.BENDY_OK:
            msg = (self._L.bendy_last_error(self._h) or b"").decode()
            raise (LinkPanic if rc == -2 else BendyError)(rc, msg)

    def clone(self) -> "Solver":  # #[derive(Clone)] solver.rs:19
        h = self._L.bendy_clone(self._h)
        if not h:
            self._ck(-3)
        c = Solver(_handle=h)
        c.gravity, c.bounds = self.gravity.copy(), Bounds(self.bounds.pos.copy(), self.bounds.size.copy())
        c.bounds_active = self.bounds_active
        return c

    # ---- snapshot (raw SoA snapshot / restore) ----------------------------------
    def save_snapshot(self, path: str):
        """The state `clone()` copies, as one flat file (format: bendy2d_b200/snapshot.py)."""
        self._ck(self._L.bendy_save_snapshot(self._h, os.fsencode(path)))

    @staticmethod
    def load_snapshot(path: str, device: int = -1) -> "Solver":
        """A new Solver that continues bit-identically to the one that was saved.  `gravity` and
        `bounds` (per-call arguments of the C ABI) come back from the last update, if one ran."""
        L = _lib.lib()
        h = L.bendy_load_snapshot(os.fsencode(path), device)
        if not h:
            raise BendyError(-1, (L.bendy_last_error(None) or b"").decode() or "bendy_load_snapshot failed")
        s = Solver(_handle=h)
        last = np.zeros(7, np.float32)
        valid = C.c_int(0)
        s._ck(L.bendy_get_last_update_args(h, _fp(last), C.byref(valid)))
        if valid.value:
            s.gravity = last[1:3].copy()
            s.bounds = Bounds(last[3:5].copy(), last[5:7].copy())
        return s

    # ---- add_* (solver.rs:52-67) -------------------------------------------------------------
    def add_particle(self, pos):
        self.add_particles(np.asarray(pos, np.float32).reshape(1, 2))

    def add_particles(self, pos_xy):
        p = _f(pos_xy).reshape(-1, 2)
        self._ck(self._L.bendy_add_particles(self._h, _fp(p), len(p)))

    def add_circle(self, circle: Circle):
        self.add_circles(circle.point.pos.reshape(1, 2), [circle.radius], circle.point.prev_pos.reshape(1, 2),
                         circle.point.acc.reshape(1, 2))

    def add_circles(self, pos_xy, radius, prev_xy=None, acc_xy=None):
        p = _f(pos_xy).reshape(-1, 2)
        q = None if prev_xy is None else _f(prev_xy).reshape(-1, 2)
        a = None if acc_xy is None else _f(acc_xy).reshape(-1, 2)
        r = _f(radius).reshape(-1)
        self._ck(self._L.bendy_add_circles(self._h, _fp(p), _fp(q), _fp(a), _fp(r), len(p)))

    def add_polygon(self, polygon: Polygon):
        pos = _f([p.pos for p in polygon.particles]).reshape(-1, 2)
        prev = _f([p.prev_pos for p in polygon.particles]).reshape(-1, 2)
        acc = _f([p.acc for p in polygon.particles]).reshape(-1, 2)
        ab = _u([[l.link.particle_a, l.link.particle_b] for l in polygon.particle_links]).reshape(-1, 2)
        ln = _f([l.link.target_distance for l in polygon.particle_links]).reshape(-1)
        self.add_polygon_raw(pos, ab, ln, polygon.is_static, polygon.center, prev, acc)

    def add_polygon_raw(self, pos_xy, link_ab, link_len, is_static, center, prev_xy=None, acc_xy=None):
        pos = _f(pos_xy).reshape(-1, 2)
        prev = None if prev_xy is None else _f(prev_xy).reshape(-1, 2)
        acc = None if acc_xy is None else _f(acc_xy).reshape(-1, 2)
        ab = _u(link_ab).reshape(-1, 2)
        ln = _f(link_len).reshape(-1)
        self._ck(self._L.bendy_add_polygon(self._h, _fp(pos), _fp(prev), _fp(acc), len(pos), _up(ab), _fp(ln),
                                           len(ln), int(bool(is_static)), float(center[0]), float(center[1])))

    def add_particle_link(self, link: ParticleLink):
        l = link.link
        self.add_particle_links([[l.particle_a, l.particle_b]], [l.target_distance])

    def add_particle_links(self, ab, lengths):
        ab64 = np.asarray(ab).reshape(-1, 2)
        if ab64.size and (ab64.min() < 0 or ab64.max() > 0xFFFFFFFF):
            raise LinkPanic(-2, "link index out of range")
        ab32, ln = _u(ab64), _f(lengths).reshape(-1)
        self._ck(self._L.bendy_add_particle_links(self._h, _up(ab32), _fp(ln), len(ln)))

    def add_circle_link(self, link: CircleLink):
        l = link.link
        self.add_circle_links([[l.particle_a, l.particle_b]], [l.target_distance])

    def add_circle_links(self, ab, lengths):
        ab32, ln = _u(np.asarray(ab).reshape(-1, 2)), _f(lengths).reshape(-1)
        self._ck(self._L.bendy_add_circle_links(self._h, _up(ab32), _fp(ln), len(ln)))

    # ---- getters (solver.rs:69-104) ------------------------------------------------------------
    def get_particle_len(self) -> int:
        return self._L.bendy_particle_len(self._h)

    def get_circles_len(self) -> int:
        return self._L.bendy_circle_len(self._h)

    def get_polygons_len(self) -> int:
        return self._L.bendy_polygon_len(self._h)

    def get_particle_links(self) -> List[ParticleLink]:
        n = self._L.bendy_particle_link_len(self._h)
        ab, ln = np.empty((n, 2), np.uint32), np.empty(n, np.float32)
        if n:
            self._ck(self._L.bendy_read_particle_links(self._h, 0, n, _up(ab), _fp(ln)))
        return [ParticleLink(Link(int(a), int(b), float(l))) for (a, b), l in zip(ab, ln)]

    def get_circle_links(self) -> List[CircleLink]:
        n = self._L.bendy_circle_link_len(self._h)
        ab, ln = np.empty((n, 2), np.uint32), np.empty(n, np.float32)
        if n:
            self._ck(self._L.bendy_read_circle_links(self._h, 0, n, _up(ab), _fp(ln)))
        return [CircleLink(Link(int(a), int(b), float(l))) for (a, b), l in zip(ab, ln)]

    def read_particles(self, first: int = 0, n: Optional[int] = None, out_pos=None, out_prev=None):
        n = self.get_particle_len() - first if n is None else n
        pos = np.empty((n, 2), np.float32) if out_pos is None else out_pos
        prev = np.empty((n, 2), np.float32) if out_prev is None else out_prev
        self._ck(self._L.bendy_read_particles(self._h, first, n, _fp(pos), _fp(prev)))
        return pos, prev

    def get_particles(self) -> ParticleArray:
        return ParticleArray(*self.read_particles())

    def get_particle(self, index: int) -> Optional[Particle]:
        if not 0 <= index < self.get_particle_len():
            return None
        pos, prev = self.read_particles(index, 1)
        return Particle(pos[0], prev[0])

    def read_circles(self):
        n = self.get_circles_len()
        pos, prev, rad = np.empty((n, 2), np.float32), np.empty((n, 2), np.float32), np.empty(n, np.float32)
        self._ck(self._L.bendy_read_circles(self._h, 0, n, _fp(pos), _fp(prev), _fp(rad)))
        return pos, prev, rad

    def get_circles(self) -> ParticleArray:
        return ParticleArray(*self.read_circles())

    def get_circle(self, index: int) -> Optional[Circle]:
        if not 0 <= index < self.get_circles_len():
            return None
        return self.get_circles()[index]

    def read_polygon(self, index: int):
        n = self._L.bendy_polygon_point_len(self._h, index)
        pos, prev, cen = np.empty((n, 2), np.float32), np.empty((n, 2), np.float32), np.empty(2, np.float32)
        st = C.c_int(0)
        self._ck(self._L.bendy_read_polygon(self._h, index, _fp(pos), _fp(prev), _fp(cen), C.byref(st)))
        return pos, prev, cen, bool(st.value)

    def get_polygon(self, index: int) -> Optional[Polygon]:
        if not 0 <= index < self.get_polygons_len():
            return None
        pos, prev, cen, st = self.read_polygon(index)
        nl = self._L.bendy_polygon_link_len(self._h, index)
        ab, ln = np.empty((nl, 2), np.uint32), np.empty(nl, np.float32)
        if nl:
            self._ck(self._L.bendy_read_polygon_links(self._h, index, _up(ab), _fp(ln)))
        return Polygon([Particle(p, q) for p, q in zip(pos, prev)],
                       [ParticleLink(Link(int(a), int(b), float(l))) for (a, b), l in zip(ab, ln)], st, cen, 1.0)

    def get_polygons(self) -> List[Polygon]:
        return [self.get_polygon(i) for i in range(self.get_polygons_len())]

    # ---- the hot path (solver.rs:106-116) --------------------------------------------------------
    def update(self, dt: float, n: int = 1):
        g, b = self.gravity, self.bounds
        self._ck(self._L.bendy_update_n(self._h, n, dt, float(g[0]), float(g[1]), float(b.pos[0]), float(b.pos[1]),
                                        float(b.size[0]), float(b.size[1])))

    def synchronize(self):
        self._ck(self._L.bendy_synchronize(self._h))

    # ---- additive API --------------------------------------------------------------------------
    def set_sub_steps(self, n: int):
        self._ck(self._L.bendy_set_sub_steps(self._h, n))

    def write_particles(self, pos_xy=None, prev_xy=None, first: int = 0):
        p = None if pos_xy is None else _f(pos_xy).reshape(-1, 2)
        q = None if prev_xy is None else _f(prev_xy).reshape(-1, 2)
        n = len(p) if p is not None else len(q)
        self._ck(self._L.bendy_write_particles(self._h, first, n, _fp(p), _fp(q)))

    def set_particle_radius(self, r: float):
        self._ck(self._L.bendy_set_particle_radius(self._h, r))

    def set_grid_cell(self, h: float):
        self._ck(self._L.bendy_set_grid_cell(self._h, h))

    def set_polygon_contact(self, on: bool):
        self._ck(self._L.bendy_set_polygon_contact(self._h, int(bool(on))))

    def set_particle_inv_mass(self, k, first: int = 0):
        k = _f(k).reshape(-1)
        self._ck(self._L.bendy_set_particle_inv_mass(self._h, first, len(k), _fp(k)))

    def set_circle_inv_mass(self, k, first: int = 0):
        k = _f(k).reshape(-1)
        self._ck(self._L.bendy_set_circle_inv_mass(self._h, first, len(k), _fp(k)))

    def set_plan_params(self, pack_points: int = 0, max_points: int = 0):
        self._ck(self._L.bendy_set_plan_params(self._h, pack_points, max_points))

    def set_link_schedule(self, mode: str = "coloured"):
        """"coloured" (default: greedy colouring, results = the reference fed the links in `link_order()`),
        or "reference": dependency-level colours, results = the reference's own insertion-order walk
        (solver.rs:143-146) bit for bit, at the price of more colours."""
        modes = {"coloured": 0, "reference": 1}
        if mode not in modes:
            raise ValueError(f"link schedule {mode!r}: one of {sorted(modes)}")
        self._ck(self._L.bendy_set_link_schedule(self._h, modes[mode]))

    # ---- schedule export (for the oracle replay) -------------------------------------------------
    def schedule_info(self) -> dict:
        info = _lib.ScheduleInfo()
        self._ck(self._L.bendy_get_schedule_info(self._h, C.byref(info)))
        return {n: getattr(info, n) for n, _ in info._fields_ if n != "reserved"}

    def link_order(self) -> np.ndarray:
        n = self._L.bendy_particle_link_len(self._h)
        perm = np.empty(n, np.uint32)
        self._ck(self._L.bendy_get_link_order(self._h, _up(perm), n))
        return perm

    def point_rank(self) -> np.ndarray:
        n = self.get_particle_len()
        rank = np.empty(n, np.uint32)
        self._ck(self._L.bendy_get_point_rank(self._h, _up(rank), n))
        return rank

    def grid(self):
        ox, oy, ih = C.c_float(), C.c_float(), C.c_float()
        nx, ny = C.c_int(), C.c_int()
        b = self.bounds
        self._ck(self._L.bendy_get_grid(self._h, float(b.pos[0]), float(b.pos[1]), float(b.size[0]),
                                        float(b.size[1]), C.byref(ox), C.byref(oy), C.byref(ih), C.byref(nx),
                                        C.byref(ny)))
        return ox.value, oy.value, ih.value, nx.value, ny.value

    # ---- measurement -----------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        self._ck(self._L.bendy_set_profiling(self._h, int(bool(on))))

    def kernel_times(self, reset: bool = False) -> dict:
        n = len(_lib.K_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_uint64 * n)()
        self._ck(self._L.bendy_get_kernel_times(self._h, ms, cnt, n, int(reset)))
        return {k: {"ms": ms[i], "launches": int(cnt[i])} for i, k in enumerate(_lib.K_CLASSES)}

    def stats(self) -> dict:
        out = (C.c_uint64 * 8)()
        self._ck(self._L.bendy_get_stats(self._h, out, 8))
        return {"circle_pass_fallbacks": int(out[0]), "by_path_bound": int(out[1]), "by_list_overflow": int(out[2]),
                "by_scale": int(out[3]), "scan_tiles": int(out[4]), "grid_cells": int(out[5]), "narrow_reverse": int(out[6])}

    def launch_count(self) -> int:
        return int(self._L.bendy_launch_count(self._h))

    def timer_start(self):
        self._ck(self._L.bendy_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._ck(self._L.bendy_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def device(self) -> int:
        return self._L.bendy_get_device(self._h)

    def device_buffers(self):
        pos, prev, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        self._ck(self._L.bendy_get_device_buffers(self._h, C.byref(pos), C.byref(prev), C.byref(n)))
        return pos.value, prev.value, n.value


def plan_links(n_points: int, ab, pack_points: int = 0, max_points: int = 0, reference_order: bool = False):
    """Host-only link planner (no GPU needed): returns rank, perm, colour, partition, info."""
    L = _lib.lib()
    ab = _u(np.asarray(ab).reshape(-1, 2))
    n = len(ab)
    rank, perm = np.empty(n_points, np.uint32), np.empty(n, np.uint32)
    colour, part = np.empty(n, np.uint32), np.empty(n, np.uint32)
    info = _lib.ScheduleInfo()
    rc = L.bendy_plan_links_scheduled(n_points, _up(ab), n, pack_points, max_points, 1 if reference_order else 0,
                                      _up(rank), _up(perm), _up(colour), _up(part), C.byref(info))
    if rc != 0:
        msg = (L.bendy_last_error(None) or b"").decode()
        raise (LinkPanic if rc == -2 else BendyError)(rc, msg)
    return rank, perm, colour, part, {k: getattr(info, k) for k, _ in info._fields_ if k != "reserved"}
