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
    return float(np.nanmax(e)) if e.size else 0.0


def compare_state(gpu, o, scale, tol=1e-5, what=""):
    """pos AND Verlet velocity (pos-prev) within tol relative (SURVEY §4 tolerance note)."""
    gp, gq = gpu.read_particles()
    op, oq = o.particles()
    stats = {"ulp_pos": max_ulp(gp, op), "ulp_prev": max_ulp(gq, oq)}
    assert rel_err(gp, op, scale) <= tol, f"{what} particle pos: {stats}"
    assert rel_err(gp - gq, op - oq, scale) <= tol * 50, f"{what} particle velocity: {stats}"
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
