"""Spatial-strip sharding of a scene across the GPUs of one box.

Whole bodies (connected components of the link graph) are assigned to vertical strips of equal
body count by centroid x.  Every substep each strip sends the owned discs that lie inside its
neighbours' halo bands; they become read-only ghost discs of the neighbour's broadphase grid.
The Jacobi contact rule only ever moves a disc on the rank that owns it, so no corrections travel
back, and because all sums are order-free the sharded run is bit-identical to the 1-GPU run.
Circles and polygons are replicated on every strip; the fixed-point corrections a strip's own discs
collect for the Circles are summed over the strips by the library (ncclAllReduce / the group's sum)
before they are applied, so every copy stays identical to the unsharded run.  Inverse masses are not
supported in strips.

  StripSolver      one process per GPU, halo over NCCL (send/recv on the solver's stream, issued by
                   libbendy2d_b200.so inside the captured substep graph); torch.distributed only
                   carries the NCCL unique id and the timing reductions.
  LocalStripGroup  all strips in one process on one GPU (device-to-device halo copies): the
                   no-cluster stand-in used by the 1-GPU tests.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace
from typing import List, Optional

import numpy as np

from . import _lib
from .scenes import Scene
from .solver import Solver

f32 = np.float32


def body_ids(scene: Scene) -> np.ndarray:
    """Connected components of the particle-link graph (unlinked particles are their own body)."""
    n = scene.n_particles
    if scene.n_links == 0:
        return np.arange(n, dtype=np.int64)
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    ab = scene.links_ab.astype(np.int64)
    g = coo_matrix((np.ones(len(ab), np.int8), (ab[:, 0], ab[:, 1])), shape=(n, n))
    _, labels = connected_components(g, directed=False)
    return labels.astype(np.int64)


@dataclass
class StripPart:
    rank: int
    world: int
    scene: Scene              # the local scene (owned bodies only)
    global_index: np.ndarray  # local particle -> index in the full scene
    x_left: float             # strip edges (-inf / +inf at the ends of the chain)
    x_right: float
    band: float
    ghost_cap: int
    stray_margin_left: Optional[float] = None   # how far an owned disc may sit inside the left / right neighbour
    stray_margin_right: Optional[float] = None  # before the ownership counts as stale (None = band / 2)
    cross: Optional[dict] = None                # links across the strip's edges (cut bodies): tables for
                                                # bendy_strip_set_cross_links + "global_links" (scene link indices)
    local_links: Optional[np.ndarray] = None    # scene link index of every link of the local scene

    @property
    def send_left_below(self) -> float:   # owned discs with x below this go to the left neighbour
        return float(self.x_left + self.band) if np.isfinite(self.x_left) else float("-inf")

    @property
    def send_right_above(self) -> float:
        return float(self.x_right - self.band) if np.isfinite(self.x_right) else float("inf")

    # an owned disc further than the stray margin (band/2, less next to a narrow strip: see partition_scene)
    # inside a neighbour's strip makes the ownership stale: as long as nothing moves more than the margin
    # - 2r between two checks, no contact can have been missed
    @property
    def stray_left(self) -> float:
        m = self.stray_margin_left if self.stray_margin_left is not None else 0.5 * self.band
        return float(self.x_left - m) if np.isfinite(self.x_left) else float("-inf")

    @property
    def stray_right(self) -> float:
        m = self.stray_margin_right if self.stray_margin_right is not None else 0.5 * self.band
        return float(self.x_right + m) if np.isfinite(self.x_right) else float("inf")


def cross_link_colours(ab: np.ndarray) -> np.ndarray:
    """Greedy colouring of the links across strip edges, in scene order (deterministic: every rank computes the
    same colours from the same scene): links of one colour share no particle."""
    used = {}
    colour = np.empty(len(ab), np.int64)
    for k, (a, b) in enumerate(ab.tolist()):
        ua, ub = used.setdefault(a, set()), used.setdefault(b, set())
        c = 0
        while c in ua or c in ub:
            c += 1
        ua.add(c), ub.add(c)
        colour[k] = c
    return colour


def partition_scene(scene: Scene, world: int, band: Optional[float] = None, bodies: Optional[np.ndarray] = None,
                    cap_factor: float = 2.0, cut_bodies: bool = False) -> List[StripPart]:
    """Split `scene` into `world` strips.  Polygons and Circles are replicated: every strip carries all of them.
    Nothing a particle does reaches a polygon, so the copies evolve identically; the corrections a strip's own
    discs collect for a Circle are summed over all strips by the library (integer all-reduce) before they are
    applied, so the Circles' copies stay identical too.

    cut_bodies=False: whole bodies (link components) go to the strip of their centroid; no link crosses an edge.
    cut_bodies=True: every particle goes to the strip of its own x (equal particle counts); links whose ends lie in
    neighbouring strips become CROSS links, relaxed after all the strips' own links in colours of their own, the
    endpoint positions exchanged before every colour (bendy_strip_set_cross_links).  For bodies wider than a strip."""

    n = scene.n_particles
    if cut_bodies:
        bodies = np.arange(n, dtype=np.int64)  # every particle is its own "body" for the geometry of the cut
    elif bodies is None:
        bodies = scene.body_of if scene.body_of is not None else body_ids(scene)
    bodies = np.asarray(bodies, np.int64)
    nb = int(bodies.max()) + 1 if n else 0
    # non-finite positions are legal state (coincident points give NaN in the reference too, link.rs:24):
    # they take no part in the geometry of the cut; a body without any finite point sits at x = 0
    px = scene.particles[:, 0].astype(np.float64)
    fin = np.isfinite(px)
    cnt = np.bincount(bodies[fin], minlength=nb).astype(np.float64)
    cx = np.bincount(bodies[fin], weights=px[fin], minlength=nb) / np.maximum(cnt, 1)
    xmin = np.full(nb, np.inf)
    xmax = np.full(nb, -np.inf)
    np.minimum.at(xmin, bodies[fin], px[fin])
    np.maximum.at(xmax, bodies[fin], px[fin])
    xmin[cnt == 0] = xmax[cnt == 0] = 0.0
    order = np.argsort(cx, kind="stable")
    # equal body counts per strip; edge = midway between the neighbouring groups' centroids
    cuts = [int(round(k * nb / world)) for k in range(world + 1)]
    strip_of_body = np.empty(nb, np.int64)
    edges = [-np.inf]
    for k in range(world):
        strip_of_body[order[cuts[k]:cuts[k + 1]]] = k
        if k + 1 < world:
            a = cx[order[cuts[k + 1] - 1]] if cuts[k + 1] > cuts[k] else cx[order[max(cuts[k + 1] - 1, 0)]]
            b = cx[order[min(cuts[k + 1], nb - 1)]]
            edges.append(0.5 * (a + b))
    edges.append(np.inf)
    if band is None:
        # a body reaches half its width past the edge it was assigned by; x2 because a disc may stray
        # band/2 before the ownership is rebalanced; + drift allowance + contact range
        half = 0.5 * float(np.max(xmax - xmin)) if nb else 0.0
        band = 2.0 * half + 2.0 + 2.0 * scene.particle_radius
    # The exchange only reaches the two neighbours.  A disc owned by strip k-1 may sit up to its stray margin
    # inside strip k before the ownership counts as stale, and so may one of strip k+1 from the other side; the
    # two must not be able to touch (2r) without either owner seeing the other, so next to a strip narrower than
    # band + 2r the margins shrink to half of what the narrow strip leaves: m = min(band/2, (width_k - 2r)/2).
    r2 = 2.0 * scene.particle_radius
    margin_into = [0.5 * band] * world  # margin_into[k]: how far a neighbour's disc may sit inside strip k
    for k in range(1, world - 1):
        width = edges[k + 1] - edges[k]
        margin_into[k] = min(0.5 * band, 0.5 * (width - r2))
        if not margin_into[k] > r2:
            raise ValueError(f"strip {k} of {world} is {width:.3f} wide: too narrow to keep the discs of strips {k - 1} and "
                             f"{k + 1} apart (contact range {r2:.3f}); use fewer strips")
    strip_of_particle = strip_of_body[bodies]
    parts = []
    lab = scene.links_ab.astype(np.int64) if scene.n_links else np.zeros((0, 2), np.int64)
    sa, sb = strip_of_particle[lab[:, 0]], strip_of_particle[lab[:, 1]]
    is_cross = sa != sb
    if is_cross.any() and not cut_bodies:
        raise ValueError("a link crosses strips: bodies must be whole (or partition with cut_bodies=True)")
    if (np.abs(sa - sb) > 1).any():
        raise ValueError("a link spans more than two neighbouring strips: use fewer strips")
    link_strip = np.where(is_cross, -1, sa)
    xsel = np.nonzero(is_cross)[0]                      # cross links, scene order
    xcol = cross_link_colours(lab[xsel]) if len(xsel) else np.zeros(0, np.int64)
    n_xcol = int(xcol.max()) + 1 if len(xsel) else 0
    for k in range(world):
        sel = np.nonzero(strip_of_particle == k)[0]
        remap = np.full(n, -1, np.int64)
        remap[sel] = np.arange(len(sel))
        lsel = np.nonzero(link_strip == k)[0]
        ab = remap[lab[lsel]]
        local = replace(scene, name=f"{scene.name} [strip {k}/{world}]", particles=scene.particles[sel].copy(),
                        links_ab=ab.astype(np.uint32), links_len=scene.links_len[lsel].copy(), body_of=None)
        xl, xr = edges[k], edges[k + 1]
        px = local.particles[:, 0]
        in_band = 0
        if np.isfinite(xl):
            in_band = max(in_band, int((px < xl + band).sum()))
        if np.isfinite(xr):
            in_band = max(in_band, int((px > xr - band).sum()))
        cross = None
        if len(xsel):
            # my share of the cross links, colour-major (scene order inside a colour)
            mine_a, mine_b = sa[xsel] == k, sb[xsel] == k
            my = np.nonzero(mine_a | mine_b)[0]
            my = my[np.argsort(xcol[my], kind="stable")]
            gl = xsel[my]
            i_am_a = mine_a[my]
            me = np.where(i_am_a, lab[gl, 0], lab[gl, 1])       # global index of my endpoint
            other = np.where(i_am_a, lab[gl, 1], lab[gl, 0])
            other_strip = strip_of_particle[other]
            # what I receive from a neighbour = the sorted set of ITS endpoints of our common links; what I send to
            # it = the sorted set of MINE (the neighbour computes the same two sets with the roles swapped)
            recv = {side: np.unique(other[other_strip == k + d]) for side, d in ((0, -1), (1, +1))}
            send = {side: np.unique(me[other_strip == k + d]) for side, d in ((0, -1), (1, +1))}
            slot = np.where(other_strip == k - 1, np.searchsorted(recv[0], other),
                            len(recv[0]) + np.searchsorted(recv[1], other))
            cs = np.searchsorted(xcol[my], np.arange(n_xcol + 1))
            cross = dict(mine=remap[me].astype(np.uint32), slot=slot.astype(np.uint32), i_am_a=i_am_a.astype(np.uint8),
                         len=scene.links_len[gl].astype(f32), colour_start=cs.astype(np.uint32), n_colours=n_xcol,
                         send_left=remap[send[0]].astype(np.uint32), send_right=remap[send[1]].astype(np.uint32),
                         n_recv_left=len(recv[0]), n_recv_right=len(recv[1]), global_links=gl)
        parts.append(StripPart(k, world, local, sel, float(xl), float(xr), float(band), 0 if world == 1 else in_band,
                               float(margin_into[k - 1]) if k > 0 else None,
                               float(margin_into[k + 1]) if k + 1 < world else None, cross, lsel))
    # both ends of an exchange use the same message size: one capacity for the whole chain
    cap = int(max(p.ghost_cap for p in parts) * cap_factor) + 1024 if world > 1 else 0
    for p in parts:
        p.ghost_cap = cap
    return parts


def sequential_link_order(parts: List[StripPart], local_orders: List[np.ndarray]) -> np.ndarray:
    """The sequential walk over the SCENE's links that the sharded run equals arithmetically:
    [strip 0's links in its solver's schedule order][strip 1's] ... [cross links, colour-major, scene order inside
    a colour].  `local_orders[k]` = strip k's Solver.link_order() (indices into its local link list).  Feed it to the
    oracle (set_link_order) to replay a sharded run on one CPU thread."""
    out = []
    for p, lo in zip(parts, local_orders):
        if p.local_links is not None and len(p.local_links):
            out.append(np.asarray(p.local_links, np.int64)[np.asarray(lo, np.int64)])
    seen = {}
    for p in parts:
        if p.cross is not None:
            cs = p.cross["colour_start"]
            for c in range(p.cross["n_colours"]):
                seen.setdefault(c, set()).update(int(g) for g in p.cross["global_links"][cs[c]:cs[c + 1]])
    for c in sorted(seen):
        out.append(np.array(sorted(seen[c]), np.int64))
    return np.concatenate(out).astype(np.uint32) if out else np.zeros(0, np.uint32)


def replicated_state(sv: Solver):
    """The state of the replicated bodies (Circles, polygons) of a strip solver, exactly as it is now: what
    rebalance() hands to the re-partitioned solvers so that they carry on bit for bit."""
    circles = sv.read_circles() if sv.get_circles_len() else None  # pos, prev, radius
    polys = []
    for k in range(sv.get_polygons_len()):
        pos, prev, cen, st = sv.read_polygon(k)
        nl = sv._L.bendy_polygon_link_len(sv._h, k)
        ab, ln = np.empty((nl, 2), np.uint32), np.empty(nl, f32)
        if nl:
            sv._ck(sv._L.bendy_read_polygon_links(sv._h, k, ab.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                  ln.ctypes.data_as(C.POINTER(C.c_float))))
        polys.append((pos, prev, cen, st, ab, ln))
    return circles, polys


def _load_part(part: StripPart, device: int, replicated=None, link_schedule: str = "coloured") -> Solver:
    """`replicated` (from replicated_state) replaces the scene's initial Circles / polygons by their current state"""
    sv = Solver(device)
    if link_schedule != "coloured":
        sv.set_link_schedule(link_schedule)
    if replicated is None:
        part.scene.load_into(sv)
    else:
        replace(part.scene, circles_pos=np.zeros((0, 2), f32), circles_r=np.zeros((0,), f32), polygons=[],
                polygons_static=[]).load_into(sv)
        circles, polys = replicated
        if circles is not None:
            sv.add_circles(circles[0], circles[2], prev_xy=circles[1])
        for pos, prev, cen, st, ab, ln in polys:
            sv.add_polygon_raw(pos, ab, ln, st, cen, prev_xy=prev)
        sv.set_polygon_contact(part.scene.polygon_contact)
    if part.cross is not None:
        x = part.cross
        up = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_uint32))
        sv._ck(sv._L.bendy_strip_set_cross_links(
            sv._h, len(x["mine"]), up(x["mine"]), up(x["slot"]),
            np.ascontiguousarray(x["i_am_a"]).ctypes.data_as(C.POINTER(C.c_uint8)),
            np.ascontiguousarray(x["len"]).ctypes.data_as(C.POINTER(C.c_float)), x["n_colours"], up(x["colour_start"]),
            len(x["send_left"]), up(x["send_left"]), len(x["send_right"]), up(x["send_right"]), x["n_recv_left"],
            x["n_recv_right"]))
    if part.world > 1:
        sv._ck(sv._L.bendy_halo_configure(sv._h, part.ghost_cap, part.send_left_below, part.send_right_above,
                                          part.stray_left, part.stray_right))
        # cells only where owned discs (up to the stray limit) and ghosts (one band beyond the edge) can be.  The
        # end strips have no edge on their open side: there the window ends 1.5 bands beyond the owned discs (a disc
        # that leaves the window is clamped into the border cells, which stays correct) - a window up to the world's
        # wall made the end ranks scan several times the cells of the interior ranks (measured: +25 us per substep)
        px = part.scene.particles[:, 0]
        px = px[np.isfinite(px)]
        lo = float(px.min()) if len(px) else 0.0
        hi = float(px.max()) if len(px) else 0.0
        x0 = part.x_left - 1.5 * part.band if np.isfinite(part.x_left) else lo - 1.5 * part.band
        x1 = part.x_right + 1.5 * part.band if np.isfinite(part.x_right) else hi + 1.5 * part.band
        sv._ck(sv._L.bendy_set_grid_window(sv._h, x0, x1))
    return sv


def _halo_stats(sv: Solver):
    a, b, o, st = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    sv._ck(sv._L.bendy_halo_stats(sv._h, C.byref(a), C.byref(b), C.byref(o), C.byref(st)))
    return a.value, b.value, o.value, st.value


class HaloError(RuntimeError):
    pass


class _StripBase:
    """Solver-like surface shared by both transports (what bench.py and the tests call)."""

    solver: Solver
    part: StripPart

    def local_scene(self) -> Scene:
        return self.part.scene

    def halo_stats(self):
        """(sent_left, sent_right, overflow, strayed) of the last substep"""
        return _halo_stats(self.solver)


def gather_global_state(dist, world: int, device_index: int, global_index: np.ndarray, pos: np.ndarray,
                        prev: np.ndarray, n_total: int):
    """Every rank contributes the (pos, prev) of the particles it owns and receives the whole scene in
    USER order.  Host-side and rare (rebalancing); float32 values travel bit-exactly as raw int32 words.
    Works on whatever the process group's backend moves (CUDA tensors under nccl, CPU tensors under gloo)."""
    import torch

    dev = f"cuda:{device_index}" if world > 1 and dist.get_backend() == "nccl" else "cpu"
    m_local = len(global_index)
    cnt = torch.tensor([m_local], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    if world > 1:
        dist.all_gather(cnts, cnt)
    else:
        cnts = [cnt]
    m = int(max(c.item() for c in cnts))
    pack = torch.zeros((m, 6), dtype=torch.int32, device=dev)
    words = np.empty((m_local, 6), np.int32)
    gi = np.asarray(global_index, np.int64)
    words[:, 0] = (gi & 0x7FFFFFFF).astype(np.int32)
    words[:, 1] = (gi >> 31).astype(np.int32)
    words[:, 2:4] = np.ascontiguousarray(pos, f32).view(np.int32).reshape(-1, 2)
    words[:, 4:6] = np.ascontiguousarray(prev, f32).view(np.int32).reshape(-1, 2)
    pack[:m_local] = torch.from_numpy(words).to(dev)
    packs = [torch.zeros_like(pack) for _ in range(world)]
    if world > 1:
        dist.all_gather(packs, pack)
    else:
        packs = [pack]
    gpos, gprev = np.empty((n_total, 2), f32), np.empty((n_total, 2), f32)
    seen = np.zeros(n_total, np.int32)
    for c, pk in zip(cnts, packs):
        a = pk[: int(c.item())].cpu().numpy()
        idx = a[:, 0].astype(np.int64) | (a[:, 1].astype(np.int64) << 31)
        gpos[idx] = np.ascontiguousarray(a[:, 2:4]).view(f32)
        gprev[idx] = np.ascontiguousarray(a[:, 4:6]).view(f32)
        seen[idx] += 1
    if not (seen == 1).all():
        raise HaloError(f"rebalance: {int((seen == 0).sum())} particles owned by no rank, "
                        f"{int((seen > 1).sum())} owned by several")
    return gpos, gprev


class StripSolver(_StripBase):
    """One strip per process / GPU; halo over NCCL issued by the C library on its own stream."""

    def __init__(self, scene: Scene, rank: int, world: int, device: int, dist=None, band: Optional[float] = None,
                 bodies: Optional[np.ndarray] = None, cut_bodies: bool = False, link_schedule: str = "coloured"):
        self.rank, self.world, self.dist, self.device_index = rank, world, dist, device
        self.full_scene, self.band = scene, band
        self.cut_bodies, self.link_schedule = cut_bodies, link_schedule
        self.bodies = bodies if bodies is not None else scene.body_of
        self._uid = None
        self._build(scene, None)

    def _build(self, scene: Scene, prev: Optional[np.ndarray], replicated=None):
        import torch

        rank, world, dist, device = self.rank, self.world, self.dist, self.device_index
        self.part = partition_scene(scene, world, self.band, self.bodies, cut_bodies=self.cut_bodies)[rank]
        self.solver = _load_part(self.part, device, replicated, self.link_schedule)
        if prev is not None:
            self.solver.write_particles(prev_xy=prev[self.part.global_index])
        if world > 1:
            L = self.solver._L
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (C.c_ubyte * 128)()
                if L.bendy_nccl_unique_id(buf) != 0:
                    raise RuntimeError((L.bendy_last_error(None) or b"").decode())
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            uid = uid.cuda(device) if dist.get_backend() == "nccl" else uid
            dist.broadcast(uid, src=0)
            raw = bytes(uid.cpu().tolist())
            self.solver._ck(L.bendy_halo_comm_nccl(self.solver._h, raw, rank, world))

    def set_particle_inv_mass(self, k):
        """ext: inverse-mass scale per particle of the FULL scene (user order); every rank keeps its own particles'.
        A ghost carries its owner's scale: the scales of the packed discs travel with their positions."""
        k = np.ascontiguousarray(k, f32).reshape(-1)
        if len(k) != self.full_scene.n_particles:
            raise ValueError("set_particle_inv_mass: one scale per particle of the full scene")
        self._inv_mass = k
        self.solver.set_particle_inv_mass(k[self.part.global_index])

    # ---- Solver surface
    def update(self, dt: float, n: int = 1):
        self.solver.update(dt, n)

    def __getattr__(self, name):  # everything else is the local solver's
        return getattr(self.solver, name)

    def check_halo(self):
        """Raises if the halo of this rank was invalid at any substep since the last check."""
        l, r, o, st = self.halo_stats()
        if o:
            raise HaloError(f"halo overflow on rank {self.rank}: ghost_cap {self.part.ghost_cap}, sent {l}/{r}")
        if st:
            raise HaloError(f"rank {self.rank}: an owned disc strayed more than band/2 = {0.5 * self.part.band:.3f} "
                            "into a neighbour's strip; call rebalance() more often")
        return l, r

    def needs_rebalance(self) -> bool:
        import torch

        on_gpu = self.world > 1 and self.dist.get_backend() == "nccl"
        flag = torch.tensor([1 if self.halo_stats()[3] else 0], device=f"cuda:{self.device_index}" if on_gpu else "cpu")
        if self.world > 1:
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MAX)
        return bool(flag.item())

    def rebalance(self):
        """Re-partition by the CURRENT positions: every rank gathers the full state (rare, host side),
        cuts new strips of equal body count and rebuilds its local solver; pos and prev travel
        bit-exactly, so the trajectory is unchanged."""
        pos, prev = self.solver.read_particles()
        gpos, gprev = gather_global_state(self.dist, self.world, self.device_index, self.part.global_index, pos, prev,
                                          self.full_scene.n_particles)
        # the replicated Circles / polygons are identical on every rank: each keeps its own copy's state
        rep = replicated_state(self.solver) if (self.full_scene.polygons or len(self.full_scene.circles_r)) else None
        self.solver = None
        self._build(replace(self.full_scene, particles=gpos), gprev, rep)


class LocalStripGroup:
    """All strips of a scene in ONE process on ONE GPU, stepped in lock-step with device-to-device
    halo copies (bendy_update_group).  Arithmetic and kernels are those of the multi-GPU path."""

    def __init__(self, scene: Scene, n_strips: int, device: int = -1, band: Optional[float] = None,
                 bodies: Optional[np.ndarray] = None):
        self.scene, self.n_strips, self.device, self.band = scene, n_strips, device, band
        self.bodies = bodies if bodies is not None else scene.body_of
        self._build(scene, None)

    def _build(self, scene: Scene, prev: Optional[np.ndarray], replicated=None):
        self.parts = partition_scene(scene, self.n_strips, self.band, self.bodies)
        self.solvers = [_load_part(p, self.device, replicated) for p in self.parts]
        if prev is not None:
            for p, s in zip(self.parts, self.solvers):
                s.write_particles(prev_xy=prev[p.global_index])
        L = self.solvers[0]._L
        for a, b in zip(self.solvers[:-1], self.solvers[1:]):
            if L.bendy_halo_connect_local(a._h, b._h) != 0:
                raise RuntimeError((L.bendy_last_error(None) or b"").decode())
        self._handles = (C.c_void_p * len(self.solvers))(*[s._h for s in self.solvers])

    def set_particle_inv_mass(self, k):
        """ext: inverse-mass scale per particle of the full scene (user order)"""
        k = np.ascontiguousarray(k, f32).reshape(-1)
        if len(k) != self.scene.n_particles:
            raise ValueError("set_particle_inv_mass: one scale per particle of the scene")
        self._inv_mass = k
        for p, s in zip(self.parts, self.solvers):
            s.set_particle_inv_mass(k[p.global_index])

    def update(self, dt: float, n: int = 1):
        s0 = self.solvers[0]
        g, b = s0.gravity, s0.bounds
        rc = s0._L.bendy_update_group(self._handles, len(self.solvers), n, dt, float(g[0]), float(g[1]),
                                      float(b.pos[0]), float(b.pos[1]), float(b.size[0]), float(b.size[1]))
        if rc != 0:
            for s in self.solvers:
                s._ck(rc)

    def read_particles(self):
        pos = np.empty((self.scene.n_particles, 2), f32)
        prev = np.empty_like(pos)
        for p, s in zip(self.parts, self.solvers):
            lp, lq = s.read_particles()
            pos[p.global_index], prev[p.global_index] = lp, lq
        return pos, prev

    def halo_stats(self):
        return [_halo_stats(s) for s in self.solvers]

    def needs_rebalance(self) -> bool:
        return any(st for _, _, _, st in self.halo_stats())

    def rebalance(self):
        pos, prev = self.read_particles()
        rep = replicated_state(self.solvers[0]) if (self.scene.polygons or len(self.scene.circles_r)) else None
        self.solvers = []
        self._build(replace(self.scene, particles=pos), prev, rep)
