"""Shared test helpers: load one Scene into the GPU solver and into the CPU oracle, compare."""
import numpy as np

from bendy2d_b200 import scenes
from oracle import bo

f32 = np.float32


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def oracle_from_scene(sc: scenes.Scene) -> bo.OracleSolver:
    o = bo.OracleSolver()
    o.set_gravity(*sc.gravity)
    o.set_bounds(*sc.bounds)
    if sc.n_particles:
        o.add_particles(sc.particles)
    if sc.n_links:
        o.add_particle_links(sc.links_ab, sc.links_len)
    for p, r in zip(sc.circles_pos, sc.circles_r):
        o.add_circle(p, float(r))
    for pts, st in zip(sc.polygons, sc.polygons_static):
        ab, ln, cen = scenes.polygon_new_tables(pts)
        o.add_polygon(pts, ab, ln, st, cen)
    o.set_sub_steps(sc.sub_steps)
    o.set_particle_radius(sc.particle_radius)
    o.set_polygon_contact(sc.polygon_contact)
    return o


def oracle_from_snapshot(snap, gravity=None, bounds=None) -> bo.OracleSolver:
    """Replay a bendy2d_b200.snapshot.Snapshot into the CPU oracle (gravity / bounds default to the
    arguments of the last update stored in the snapshot)."""
    o = bo.OracleSolver()
    last = snap.last_update
    o.set_gravity(*(gravity if gravity is not None else last[1:3]))
    o.set_bounds(*(bounds if bounds is not None else last[3:7]))
    if len(snap.particles_pos):
        o.add_particles(snap.particles_pos)
        o.write_particles(snap.particles_pos, snap.particles_prev)
    if len(snap.particle_links_len):
        o.add_particle_links(snap.particle_links_ab, snap.particle_links_len)
    for p, q, a, r in zip(snap.circles_pos, snap.circles_prev, snap.circles_acc, snap.circles_radius):
        o.add_circle(p, float(r), q, (float(a[0]), float(a[1])))
    for a, b, l in zip(snap.circle_links_ab[:, 0], snap.circle_links_ab[:, 1], snap.circle_links_len):
        o.add_circle_link(int(a), int(b), float(l))
    for k in range(snap.n_polygons):
        pg = snap.polygon(k)
        o.add_polygon(pg["pos"], pg["link_ab"], pg["link_len"], pg["is_static"], pg["center"], pg["prev"], pg["acc"])
    o.set_sub_steps(snap.sub_steps)
    o.set_particle_radius(snap.particle_radius)
    o.set_polygon_contact(snap.polygon_contact)
    if snap.particles_inv_mass is not None:
        o.set_particle_inv_mass(0, snap.particles_inv_mass)
    if snap.circles_inv_mass is not None:
        o.set_circle_inv_mass(0, snap.circles_inv_mass)
    return o


def snapshot_from_scene(sc: scenes.Scene):
    """The snapshot a freshly loaded Scene would produce (no update yet)."""
    from bendy2d_b200.snapshot import Snapshot

    s = Snapshot(sub_steps=sc.sub_steps, particle_radius=sc.particle_radius, polygon_contact=sc.polygon_contact)
    s.particles_pos = np.ascontiguousarray(sc.particles, f32).reshape(-1, 2)
    s.particles_prev = s.particles_pos.copy()
    s.particle_links_ab = np.ascontiguousarray(sc.links_ab, np.uint32).reshape(-1, 2)
    s.particle_links_len = np.ascontiguousarray(sc.links_len, f32)
    s.circles_pos = np.ascontiguousarray(sc.circles_pos, f32).reshape(-1, 2)
    s.circles_prev = s.circles_pos.copy()
    s.circles_acc = np.zeros_like(s.circles_pos)
    s.circles_radius = np.ascontiguousarray(sc.circles_r, f32)
    starts, nvs, lstarts, nls, cens, pts_all, ab_all, ln_all = [], [], [], [], [], [], [], []
    p0 = l0 = 0
    for pts in sc.polygons:
        ab, ln, cen = scenes.polygon_new_tables(pts)
        starts.append(p0), nvs.append(len(pts)), lstarts.append(l0), nls.append(len(ln)), cens.append(cen)
        pts_all.append(np.asarray(pts, f32)), ab_all.append(np.asarray(ab, np.uint32).reshape(-1, 2)), ln_all.append(ln)
        p0, l0 = p0 + len(pts), l0 + len(ln)
    if starts:
        s.poly_start, s.poly_nv = np.array(starts, np.uint32), np.array(nvs, np.uint32)
        s.poly_link_start, s.poly_nl = np.array(lstarts, np.uint32), np.array(nls, np.uint32)
        s.poly_static = np.array(sc.polygons_static, bool)
        s.poly_center = np.array(cens, f32).reshape(-1, 2)
        s.poly_points_pos = np.concatenate(pts_all).astype(f32)
        s.poly_points_prev = s.poly_points_pos.copy()
        s.poly_points_acc = np.zeros_like(s.poly_points_pos)
        s.poly_links_ab = np.concatenate(ab_all).astype(np.uint32)
        s.poly_links_len = np.concatenate(ln_all).astype(f32)
    return s


def sync_schedule(gpu, o: bo.OracleSolver, sc: scenes.Scene):
    """Feed the oracle the GPU's schedule: link order, in-cell rank, grid."""
    if sc.n_links:
        o.set_link_order(gpu.link_order())
    if sc.particle_radius > 0 and sc.n_particles:
        o.set_point_rank(gpu.point_rank())
        o.set_grid(*gpu.grid())


def max_ulp(a, b):
    a, b = np.ascontiguousarray(a, f32), np.ascontiguousarray(b, f32)
    same_nan = np.isnan(a) & np.isnan(b)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)
    d[same_nan] = 0
    return int(d.max()) if d.size else 0


def rel_err(got, ref, scale):
    """max |got-ref| / max(|ref|, scale*eps_f32) — the SURVEY §4 tolerance definition."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    den = np.maximum(np.abs(ref), scale * np.finfo(np.float32).eps)
    with np.errstate(invalid="ignore"):
        e = np.abs(got - ref) / den
    e[np.isnan(got) & np.isnan(ref)] = 0.0
    e[got == ref] = 0.0  # equal infinities (inf - inf is NaN above)
    return float(np.nanmax(e)) if e.size else 0.0


def compare_state(gpu, o, scale, tol=1e-5, what=""):
    """pos AND Verlet velocity (pos-prev) within tol relative (SURVEY §4 tolerance note)."""
    gp, gq = gpu.read_particles()
    op, oq = o.particles()
    stats = {"ulp_pos": max_ulp(gp, op), "ulp_prev": max_ulp(gq, oq)}
    assert rel_err(gp, op, scale) <= tol, f"{what} particle pos: {stats}"
    assert rel_err(gp - gq, op - oq, scale) <= tol, f"{what} particle velocity: {stats}"
    if o.circle_len():
        cp, cq, _ = gpu.read_circles()
        ocp, ocq, _ = o.circles()
        stats["ulp_circle"] = max_ulp(cp, ocp)
        assert rel_err(cp, ocp, scale) <= tol, f"{what} circle pos: {stats}"
        assert rel_err(cq, ocq, scale) <= tol, f"{what} circle prev: {stats}"
    for k in range(o.polygon_len()):
        pp, pq, pc, _ = gpu.read_polygon(k)
        op_, oq_, oc_ = o.polygon(k)
        assert rel_err(pp, op_, scale) <= tol, f"{what} polygon {k} pos"
        assert rel_err(pq, oq_, scale) <= tol, f"{what} polygon {k} prev"
        assert rel_err(pc, oc_, scale) <= tol, f"{what} polygon {k} centre"
    return stats
