"""ctypes binding of the CPU oracle (oracle/bendy_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (bendy2d_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libbendy_oracle.so")

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (oracle/Makefile). Returns the .so path."""
    src = os.path.join(_HERE, "bendy_oracle.cpp")
    hdr = os.path.join(_HERE, "bendy_oracle.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    sz = C.c_size_t
    fl = C.c_float
    sig = {
        "bo_create": (vp, []),
        "bo_destroy": (None, [vp]),
        "bo_clone": (vp, [vp]),
        "bo_set_gravity": (None, [vp, fl, fl]),
        "bo_set_bounds": (None, [vp, fl, fl, fl, fl]),
        "bo_add_particle": (None, [vp, fl, fl]),
        "bo_add_particles": (None, [vp, f32p, sz]),
        "bo_add_particle_links": (None, [vp, u32p, f32p, sz]),
        "bo_add_circle": (None, [vp, fl, fl, fl, fl, fl, fl, fl]),
        "bo_add_polygon": (C.c_int, [vp, f32p, f32p, f32p, sz, u32p, f32p, sz, C.c_int, fl, fl]),
        "bo_add_polygon_new": (C.c_int, [vp, f32p, sz, C.c_int]),
        "bo_add_polygon_circle": (C.c_int, [vp, fl, fl, fl, sz, C.c_int]),
        "bo_add_particle_link": (None, [vp, sz, sz, fl]),
        "bo_add_circle_link": (None, [vp, sz, sz, fl]),
        "bo_update": (C.c_int, [vp, fl]),
        "bo_particle_len": (sz, [vp]),
        "bo_circle_len": (sz, [vp]),
        "bo_polygon_len": (sz, [vp]),
        "bo_particle_link_len": (sz, [vp]),
        "bo_read_particles": (None, [vp, f32p, f32p]),
        "bo_write_particles": (None, [vp, f32p, f32p]),
        "bo_read_circles": (None, [vp, f32p, f32p, f32p]),
        "bo_polygon_point_len": (sz, [vp, sz]),
        "bo_polygon_link_len": (sz, [vp, sz]),
        "bo_read_polygon": (None, [vp, sz, f32p, f32p, f32p]),
        "bo_read_polygon_links": (None, [vp, sz, u32p, f32p]),
        "bo_set_link_order": (C.c_int, [vp, u32p, sz]),
        "bo_set_sub_steps": (None, [vp, C.c_uint16]),
        "bo_ext_set_particle_radius": (None, [vp, fl]),
        "bo_ext_set_grid": (None, [vp, fl, fl, fl, C.c_int, C.c_int]),
        "bo_ext_set_point_rank": (None, [vp, u32p, sz]),
        "bo_ext_set_particle_inv_mass": (None, [vp, sz, sz, f32p]),
        "bo_ext_set_circle_inv_mass": (None, [vp, sz, sz, f32p]),
        "bo_ext_set_polygon_contact": (None, [vp, C.c_int]),
        "bo_prim_particle_update": (None, [f32p, f32p, f32p, fl]),
        "bo_prim_particle_bounds": (None, [f32p, f32p, fl, fl, fl, fl]),
        "bo_prim_circle_bounds": (None, [f32p, f32p, fl, fl, fl, fl, fl]),
        "bo_prim_link_solve": (None, [f32p, f32p, fl]),
        "bo_prim_circle_link_solve": (None, [f32p, f32p, fl, fl, fl]),
        "bo_prim_circle_solve": (C.c_int, [f32p, f32p, fl, fl]),
        "bo_prim_line_intersection": (C.c_int, [f32p, f32p, f32p, f32p, f32p]),
        "bo_prim_resolve_line_intersection": (C.c_int, [f32p, f32p, f32p, f32p, f32p, f32p]),
        "bo_prim_solve_polygon_single": (None, [f32p, sz, f32p, f32p, sz, f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _u(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(f32p)


def _up(a: np.ndarray):
    return a.ctypes.data_as(u32p)


class OraclePanic(RuntimeError):
    """The reference would have panicked (link index checks, link.rs:19-21)."""


class OracleSolver:
    """Mirror of the reference `Solver` (solver.rs:20-116) on the CPU oracle."""

    def __init__(self, _handle=None):
        self._L = lib()
        self._h = _handle if _handle is not None else self._L.bo_create()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.bo_destroy(h)

    def clone(self) -> "OracleSolver":
        return OracleSolver(self._L.bo_clone(self._h))

    # pub fields
    def set_gravity(self, gx, gy):
        self._L.bo_set_gravity(self._h, gx, gy)

    def set_bounds(self, bx, by, sx, sy):
        self._L.bo_set_bounds(self._h, bx, by, sx, sy)

    # add_* (solver.rs:52-67)
    def add_particle(self, x, y):
        self._L.bo_add_particle(self._h, x, y)

    def add_particles(self, pos_xy):
        p = _f(pos_xy).reshape(-1, 2)
        self._L.bo_add_particles(self._h, _fp(p), len(p))

    def add_circle(self, pos, radius, prev=None, acc=(0.0, 0.0)):
        prev = pos if prev is None else prev
        self._L.bo_add_circle(self._h, pos[0], pos[1], prev[0], prev[1], acc[0], acc[1], radius)

    def add_polygon(self, pos_xy, link_ab, link_len, is_static, center, prev_xy=None, acc_xy=None):
        pos = _f(pos_xy).reshape(-1, 2)
        prev = None if prev_xy is None else _f(prev_xy).reshape(-1, 2)
        acc = None if acc_xy is None else _f(acc_xy).reshape(-1, 2)
        ab = _u(link_ab).reshape(-1, 2)
        ln = _f(link_len).reshape(-1)
        rc = self._L.bo_add_polygon(
            self._h, _fp(pos), _fp(prev) if prev is not None else None, _fp(acc) if acc is not None else None,
            len(pos), _up(ab), _fp(ln), len(ln), int(bool(is_static)), center[0], center[1])
        if rc != 0:
            raise ValueError("bo_add_polygon failed")

    def add_polygon_new(self, pts_xy, is_static):
        p = _f(pts_xy).reshape(-1, 2)
        if self._L.bo_add_polygon_new(self._h, _fp(p), len(p), int(bool(is_static))) != 0:
            raise ValueError("bo_add_polygon_new failed")

    def add_polygon_circle(self, radius, pos, point_count, is_static):
        if self._L.bo_add_polygon_circle(self._h, radius, pos[0], pos[1], point_count, int(bool(is_static))) != 0:
            raise ValueError("bo_add_polygon_circle failed")

    def add_particle_link(self, a, b, length):
        self._L.bo_add_particle_link(self._h, a, b, length)

    def add_particle_links(self, ab, lengths):
        ab = _u(ab).reshape(-1, 2)
        ln = _f(lengths).reshape(-1)
        self._L.bo_add_particle_links(self._h, _up(ab), _fp(ln), len(ln))

    def add_circle_link(self, a, b, length):
        self._L.bo_add_circle_link(self._h, a, b, length)

    def update(self, dt):
        rc = self._L.bo_update(self._h, dt)
        if rc == -1:
            raise OraclePanic("reference would panic: invalid link indices")
        if rc != 0:
            raise ValueError(f"bo_update failed: {rc}")

    # getters
    def particle_len(self):
        return self._L.bo_particle_len(self._h)

    def circle_len(self):
        return self._L.bo_circle_len(self._h)

    def polygon_len(self):
        return self._L.bo_polygon_len(self._h)

    def particles(self):
        n = self.particle_len()
        pos = np.empty((n, 2), np.float32)
        prev = np.empty((n, 2), np.float32)
        if n:
            self._L.bo_read_particles(self._h, _fp(pos), _fp(prev))
        return pos, prev

    def write_particles(self, pos, prev):
        pos, prev = _f(pos), _f(prev)
        self._L.bo_write_particles(self._h, _fp(pos), _fp(prev))

    def circles(self):
        n = self.circle_len()
        pos = np.empty((n, 2), np.float32)
        prev = np.empty((n, 2), np.float32)
        rad = np.empty((n,), np.float32)
        if n:
            self._L.bo_read_circles(self._h, _fp(pos), _fp(prev), _fp(rad))
        return pos, prev, rad

    def polygon(self, idx):
        n = self._L.bo_polygon_point_len(self._h, idx)
        pos = np.empty((n, 2), np.float32)
        prev = np.empty((n, 2), np.float32)
        center = np.empty((2,), np.float32)
        self._L.bo_read_polygon(self._h, idx, _fp(pos), _fp(prev), _fp(center))
        return pos, prev, center

    def polygon_links(self, idx):
        n = self._L.bo_polygon_link_len(self._h, idx)
        ab = np.empty((n, 2), np.uint32)
        ln = np.empty((n,), np.float32)
        if n:
            self._L.bo_read_polygon_links(self._h, idx, _up(ab), _fp(ln))
        return ab, ln

    # replay + extensions
    def set_link_order(self, perm):
        if perm is None:
            self._L.bo_set_link_order(self._h, None, 0)
            return
        p = _u(perm)
        if self._L.bo_set_link_order(self._h, _up(p), len(p)) != 0:
            raise ValueError("link order is not a permutation of the particle links")

    def set_sub_steps(self, n):
        self._L.bo_set_sub_steps(self._h, n)

    def set_particle_radius(self, r):
        self._L.bo_ext_set_particle_radius(self._h, r)

    def set_grid(self, ox, oy, inv_h, nx, ny):
        self._L.bo_ext_set_grid(self._h, ox, oy, inv_h, nx, ny)

    def set_point_rank(self, rank):
        if rank is None:
            self._L.bo_ext_set_point_rank(self._h, None, 0)
        else:
            r = _u(rank)
            self._L.bo_ext_set_point_rank(self._h, _up(r), len(r))

    def set_particle_inv_mass(self, first, k):
        k = _f(k)
        self._L.bo_ext_set_particle_inv_mass(self._h, first, len(k), _fp(k))

    def set_circle_inv_mass(self, first, k):
        k = _f(k)
        self._L.bo_ext_set_circle_inv_mass(self._h, first, len(k), _fp(k))

    def set_polygon_contact(self, on):
        self._L.bo_ext_set_polygon_contact(self._h, int(bool(on)))


# ---- primitives (for KATs) ----
def prim_link_solve(a, b, length):
    a, b = _f(a).copy(), _f(b).copy()
    lib().bo_prim_link_solve(_fp(a), _fp(b), length)
    return a, b


def prim_circle_link_solve(a, b, ra, rb, length):
    a, b = _f(a).copy(), _f(b).copy()
    lib().bo_prim_circle_link_solve(_fp(a), _fp(b), ra, rb, length)
    return a, b


def prim_circle_solve(p1, p2, r1, r2):
    p1, p2 = _f(p1).copy(), _f(p2).copy()
    hit = lib().bo_prim_circle_solve(_fp(p1), _fp(p2), r1, r2)
    return bool(hit), p1, p2


def prim_particle_update(pos, prev, acc, dt):
    pos, prev, acc = _f(pos).copy(), _f(prev).copy(), _f(acc).copy()
    lib().bo_prim_particle_update(_fp(pos), _fp(prev), _fp(acc), dt)
    return pos, prev, acc


def prim_particle_bounds(pos, prev, bounds):
    pos, prev = _f(pos).copy(), _f(prev).copy()
    lib().bo_prim_particle_bounds(_fp(pos), _fp(prev), *bounds)
    return pos, prev


def prim_circle_bounds(pos, prev, r, bounds):
    pos, prev = _f(pos).copy(), _f(prev).copy()
    lib().bo_prim_circle_bounds(_fp(pos), _fp(prev), r, *bounds)
    return pos, prev


def prim_line_intersection(p1, p2, p3, p4):
    out = np.zeros(2, np.float32)
    hit = lib().bo_prim_line_intersection(_fp(_f(p1)), _fp(_f(p2)), _fp(_f(p3)), _fp(_f(p4)), _fp(out))
    return out if hit else None


def prim_resolve_line_intersection(a, b, q, other_center, self_center):
    out = np.zeros(6, np.float32)
    hit = lib().bo_prim_resolve_line_intersection(
        _fp(_f(a)), _fp(_f(b)), _fp(_f(q)), _fp(_f(other_center)), _fp(_f(self_center)), _fp(out))
    return out.reshape(3, 2) if hit else None


def prim_solve_polygon_single(self_xy, self_center, other_xy, other_center):
    s, o = _f(self_xy).copy().reshape(-1, 2), _f(other_xy).copy().reshape(-1, 2)
    lib().bo_prim_solve_polygon_single(_fp(s), len(s), _fp(_f(self_center)), _fp(o), len(o), _fp(_f(other_center)))
    return s, o
