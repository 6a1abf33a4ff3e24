"""The on-disk snapshot of a Solver (`bendy_save_snapshot` / `bendy_load_snapshot`), read and written
with numpy only.

The reference has no file format; `#[derive(Clone)]` on `Solver` (reference src/solver.rs:19) is its only
checkpoint.  The snapshot is that state as one flat file of little-endian 4-byte words, so a parity
failure found on a GPU box can be carried elsewhere and replayed (e.g. into the CPU oracle by the
tests).  This module never touches the GPU library: it is the format's second, independent
implementation, which is what the round-trip tests compare the C side against.

Layout (writer: bendy2d_b200/csrc/solver.cu):

    magic "B2DSNAP1"
    u32 version (=1), sub_steps | f32 particle_radius, grid_cell
    u32 flags (bit 0 polygon_contact, bit 1 reference-order link schedule), pack_points, max_points, has_last_update
    f32 last update: dt, gravity x, y, bounds x, y, w, h, 0
    u64 n_particles, n_circles, n_polygons, n_polygon_points, n_particle_links, n_circle_links,
        n_polygon_links, has_particle_inv_mass, has_circle_inv_mass
    particles: pos[2n] prev[2n] (inv_mass[n])
    circles:   pos[2n] prev[2n] acc[2n] radius[n] (inv_mass[n])
    particle links: ab[2n] u32, length[n]          (insertion order, link.rs:5-10)
    circle links:   ab[2n] u32, length[n]
    polygon table:  per polygon u32 start, nv, link_start, nl, is_static, f32 centre x, y
    polygon points: pos[2n] prev[2n] acc[2n]       (polygon-major)
    polygon links:  ab[2n] u32 (polygon-local indices, polygon.rs:218-223), length[n]
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

MAGIC = b"B2DSNAP1"
VERSION = 1
_HEADER = struct.Struct("<8sIIffIIII8f9Q")  # 144 bytes

f32 = np.float32
u32 = np.uint32


class SnapshotError(ValueError):
    pass


def _z2():
    return np.zeros((0, 2), f32)


@dataclass
class Snapshot:
    sub_steps: int = 1
    particle_radius: float = 0.0
    grid_cell: float = 0.0
    polygon_contact: bool = False
    reference_link_order: bool = False
    pack_points: int = 512
    max_points: int = 4096
    last_update: Optional[np.ndarray] = None  # dt, gx, gy, bx, by, bw, bh of the last update(), or None
    particles_pos: np.ndarray = field(default_factory=_z2)
    particles_prev: np.ndarray = field(default_factory=_z2)
    particles_inv_mass: Optional[np.ndarray] = None
    circles_pos: np.ndarray = field(default_factory=_z2)
    circles_prev: np.ndarray = field(default_factory=_z2)
    circles_acc: np.ndarray = field(default_factory=_z2)
    circles_radius: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    circles_inv_mass: Optional[np.ndarray] = None
    particle_links_ab: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), u32))
    particle_links_len: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    circle_links_ab: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), u32))
    circle_links_len: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    # polygon table: one row per polygon
    poly_start: np.ndarray = field(default_factory=lambda: np.zeros(0, u32))
    poly_nv: np.ndarray = field(default_factory=lambda: np.zeros(0, u32))
    poly_link_start: np.ndarray = field(default_factory=lambda: np.zeros(0, u32))
    poly_nl: np.ndarray = field(default_factory=lambda: np.zeros(0, u32))
    poly_static: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))
    poly_center: np.ndarray = field(default_factory=_z2)
    poly_points_pos: np.ndarray = field(default_factory=_z2)
    poly_points_prev: np.ndarray = field(default_factory=_z2)
    poly_points_acc: np.ndarray = field(default_factory=_z2)
    poly_links_ab: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), u32))
    poly_links_len: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))

    # -- convenience -----------------------------------------------------------------------------
    @property
    def n_polygons(self) -> int:
        return len(self.poly_nv)

    def polygon(self, k: int) -> dict:
        """Points, links (polygon-local indices), static flag and cached centre of polygon k."""
        p0, nv = int(self.poly_start[k]), int(self.poly_nv[k])
        l0, nl = int(self.poly_link_start[k]), int(self.poly_nl[k])
        return {
            "pos": self.poly_points_pos[p0:p0 + nv], "prev": self.poly_points_prev[p0:p0 + nv],
            "acc": self.poly_points_acc[p0:p0 + nv], "link_ab": self.poly_links_ab[l0:l0 + nl],
            "link_len": self.poly_links_len[l0:l0 + nl], "is_static": bool(self.poly_static[k]),
            "center": self.poly_center[k],
        }

    def validate(self) -> None:
        """The rules bendy_load_snapshot enforces (sizes; a < b < n for every link, link.rs:19-21)."""
        nP, nC, nG = len(self.particles_pos), len(self.circles_pos), len(self.poly_points_pos)
        if not 1 <= self.sub_steps <= 0xFFFF:
            raise SnapshotError("sub_steps out of range")
        for name, arr, n in (("particles_prev", self.particles_prev, nP), ("circles_prev", self.circles_prev, nC),
                             ("circles_acc", self.circles_acc, nC), ("circles_radius", self.circles_radius, nC),
                             ("poly_points_prev", self.poly_points_prev, nG),
                             ("poly_points_acc", self.poly_points_acc, nG)):
            if len(arr) != n:
                raise SnapshotError(f"{name} has {len(arr)} rows, expected {n}")
        if self.particles_inv_mass is not None and len(self.particles_inv_mass) != nP:
            raise SnapshotError("particles_inv_mass length")
        if self.circles_inv_mass is not None and len(self.circles_inv_mass) != nC:
            raise SnapshotError("circles_inv_mass length")
        for name, ab, ln, n in (("particle", self.particle_links_ab, self.particle_links_len, nP),
                                ("circle", self.circle_links_ab, self.circle_links_len, nC)):
            if len(ab) != len(ln):
                raise SnapshotError(f"{name} link arrays differ in length")
            if len(ab) and not (np.all(ab[:, 0] < ab[:, 1]) and np.all(ab[:, 1] < n)):
                raise SnapshotError(f"{name} link out of range")
        pts = lks = 0
        for k in range(self.n_polygons):
            nv, nl = int(self.poly_nv[k]), int(self.poly_nl[k])
            if int(self.poly_start[k]) != pts or int(self.poly_link_start[k]) != lks or nv == 0:
                raise SnapshotError("polygon table inconsistent")
            ab = self.poly_links_ab[lks:lks + nl]
            if len(ab) != nl or (nl and not (np.all(ab[:, 0] < ab[:, 1]) and np.all(ab[:, 1] < nv))):
                raise SnapshotError("polygon link out of range")
            pts, lks = pts + nv, lks + nl
        if pts != nG or lks != len(self.poly_links_len) or len(self.poly_links_ab) != lks:
            raise SnapshotError("polygon table does not cover the polygon points / links")

    # -- file i/o -------------------------------------------------------------------------------
    def to_bytes(self) -> bytes:
        self.validate()
        last = np.zeros(8, f32)
        if self.last_update is not None:
            last[:7] = np.asarray(self.last_update, f32)
        counts = (len(self.particles_pos), len(self.circles_pos), self.n_polygons, len(self.poly_points_pos),
                  len(self.particle_links_len), len(self.circle_links_len), len(self.poly_links_len),
                  int(self.particles_inv_mass is not None), int(self.circles_inv_mass is not None))
        head = _HEADER.pack(MAGIC, VERSION, self.sub_steps, self.particle_radius, self.grid_cell,
                            int(self.polygon_contact) | (2 if self.reference_link_order else 0), self.pack_points, self.max_points,
                            int(self.last_update is not None), *[float(x) for x in last], *counts)
        table = np.zeros((self.n_polygons, 7), u32)
        if self.n_polygons:
            table[:, 0], table[:, 1] = self.poly_start, self.poly_nv
            table[:, 2], table[:, 3] = self.poly_link_start, self.poly_nl
            table[:, 4] = self.poly_static.astype(u32)
            table[:, 5:7] = np.ascontiguousarray(self.poly_center, f32).view(u32)
        parts = [head]

        def put(a, dt):
            if a is not None:
                parts.append(np.ascontiguousarray(a, dt).tobytes())

        put(self.particles_pos, f32), put(self.particles_prev, f32), put(self.particles_inv_mass, f32)
        put(self.circles_pos, f32), put(self.circles_prev, f32), put(self.circles_acc, f32)
        put(self.circles_radius, f32), put(self.circles_inv_mass, f32)
        put(self.particle_links_ab, u32), put(self.particle_links_len, f32)
        put(self.circle_links_ab, u32), put(self.circle_links_len, f32)
        put(table, u32)
        put(self.poly_points_pos, f32), put(self.poly_points_prev, f32), put(self.poly_points_acc, f32)
        put(self.poly_links_ab, u32), put(self.poly_links_len, f32)
        return b"".join(parts)

    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(self.to_bytes())

    @staticmethod
    def from_bytes(buf: bytes) -> "Snapshot":
        if len(buf) < _HEADER.size or buf[:8] != MAGIC:
            raise SnapshotError("not a bendy2d snapshot (bad magic)")
        h = _HEADER.unpack_from(buf, 0)
        version, sub_steps, rp, cell, contact, pack, maxp, has_last = h[1:9]
        last, counts = np.array(h[9:17], f32), h[17:26]
        if version != VERSION:
            raise SnapshotError("unsupported snapshot version")
        nP, nC, nPoly, nG, nPL, nCL, nGL, has_pk, has_ck = counts
        if max(counts[:7]) > 0x7FFFFFF0:
            raise SnapshotError("corrupt header (counts)")
        off = _HEADER.size

        def take(n, dt, shape=None):
            nonlocal off
            nbytes = 4 * n
            if off + nbytes > len(buf):
                raise SnapshotError("truncated snapshot")
            a = np.frombuffer(buf, dt, n, off).copy()
            off += nbytes
            return a if shape is None else a.reshape(shape)

        s = Snapshot(sub_steps=sub_steps, particle_radius=rp, grid_cell=cell, polygon_contact=bool(contact & 1),
                     reference_link_order=bool(contact & 2),
                     pack_points=pack, max_points=maxp, last_update=last[:7].copy() if has_last else None)
        s.particles_pos, s.particles_prev = take(2 * nP, f32, (-1, 2)), take(2 * nP, f32, (-1, 2))
        s.particles_inv_mass = take(nP, f32) if has_pk else None
        s.circles_pos, s.circles_prev = take(2 * nC, f32, (-1, 2)), take(2 * nC, f32, (-1, 2))
        s.circles_acc, s.circles_radius = take(2 * nC, f32, (-1, 2)), take(nC, f32)
        s.circles_inv_mass = take(nC, f32) if has_ck else None
        s.particle_links_ab, s.particle_links_len = take(2 * nPL, u32, (-1, 2)), take(nPL, f32)
        s.circle_links_ab, s.circle_links_len = take(2 * nCL, u32, (-1, 2)), take(nCL, f32)
        table = take(7 * nPoly, u32, (-1, 7))
        s.poly_start, s.poly_nv = table[:, 0].copy(), table[:, 1].copy()
        s.poly_link_start, s.poly_nl = table[:, 2].copy(), table[:, 3].copy()
        s.poly_static = table[:, 4] != 0
        s.poly_center = np.ascontiguousarray(table[:, 5:7]).view(f32).reshape(-1, 2)
        s.poly_points_pos, s.poly_points_prev = take(2 * nG, f32, (-1, 2)), take(2 * nG, f32, (-1, 2))
        s.poly_points_acc = take(2 * nG, f32, (-1, 2))
        s.poly_links_ab, s.poly_links_len = take(2 * nGL, u32, (-1, 2)), take(nGL, f32)
        if off != len(buf):
            raise SnapshotError("trailing bytes after the snapshot")
        s.validate()
        return s

    @staticmethod
    def load(path: str) -> "Snapshot":
        with open(path, "rb") as f:
            return Snapshot.from_bytes(f.read())

    # -- replay ---------------------------------------------------------------------------------
    def load_into(self, solver) -> None:
        """Rebuild the scene through the public add_* surface (anything shaped like
        bendy2d_b200.Solver: the GPU solver or the tests' oracle wrapper)."""
        if len(self.particles_pos):
            solver.add_particles(self.particles_pos)
            if not np.array_equal(self.particles_pos.view(u32), self.particles_prev.view(u32)):
                solver.write_particles(None, self.particles_prev)
        if len(self.particle_links_len):
            solver.add_particle_links(self.particle_links_ab, self.particle_links_len)
        if len(self.circles_radius):
            solver.add_circles(self.circles_pos, self.circles_radius, self.circles_prev, self.circles_acc)
        if len(self.circle_links_len):
            solver.add_circle_links(self.circle_links_ab, self.circle_links_len)
        for k in range(self.n_polygons):
            g = self.polygon(k)
            solver.add_polygon_raw(g["pos"], g["link_ab"], g["link_len"], g["is_static"], g["center"], g["prev"],
                                   g["acc"])
        solver.set_sub_steps(self.sub_steps)
        solver.set_particle_radius(self.particle_radius)
        if self.grid_cell:
            solver.set_grid_cell(self.grid_cell)
        solver.set_polygon_contact(self.polygon_contact)
        if self.reference_link_order:
            solver.set_link_schedule("reference")
        if self.particles_inv_mass is not None:
            solver.set_particle_inv_mass(self.particles_inv_mass)
        if self.circles_inv_mass is not None:
            solver.set_circle_inv_mass(self.circles_inv_mass)
