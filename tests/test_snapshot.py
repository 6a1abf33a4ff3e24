"""Snapshot format: numpy reader/writer vs the C ABI's bendy_save/load_snapshot."""
import ctypes as C
import os

import numpy as np
import pytest

from bendy2d_b200 import _lib, scenes
from bendy2d_b200.snapshot import Snapshot, SnapshotError

from helpers import (bits, compare_state, oracle_from_scene, oracle_from_snapshot, snapshot_from_scene,
                     sync_schedule)

f32, u32 = np.float32, np.uint32


def small_snapshot() -> Snapshot:
    r = np.random.default_rng(7)
    s = Snapshot(sub_steps=8, particle_radius=0.1, grid_cell=0.5, polygon_contact=True,
                 last_update=np.array([1 / 60, 0.0, 98.2, 0, 0, 100, 100], f32))
    s.particles_pos = r.uniform(0, 100, (50, 2)).astype(f32)
    s.particles_prev = s.particles_pos + f32(0.01)
    s.particles_inv_mass = r.uniform(0, 2, 50).astype(f32)
    s.particle_links_ab = np.array([[0, 1], [1, 2], [3, 49]], u32)
    s.particle_links_len = np.array([1.0, 2.0, 3.5], f32)
    s.circles_pos = r.uniform(0, 100, (3, 2)).astype(f32)
    s.circles_prev = s.circles_pos.copy()
    s.circles_acc = np.array([[0, 0], [1, 2], [0, 0]], f32)
    s.circles_radius = np.array([1, 2, 3], f32)
    s.circle_links_ab = np.array([[0, 2]], u32)
    s.circle_links_len = np.array([10.0], f32)
    s.poly_start, s.poly_nv = np.array([0, 3], u32), np.array([3, 4], u32)
    s.poly_link_start, s.poly_nl = np.array([0, 3], u32), np.array([3, 2], u32)
    s.poly_static = np.array([True, False])
    s.poly_center = np.array([[1, 1], [5, 5]], f32)
    s.poly_points_pos = r.uniform(0, 10, (7, 2)).astype(f32)
    s.poly_points_prev = s.poly_points_pos.copy()
    s.poly_points_acc = np.zeros((7, 2), f32)
    s.poly_links_ab = np.array([[0, 1], [1, 2], [0, 2], [0, 1], [2, 3]], u32)
    s.poly_links_len = np.arange(5, dtype=f32) + 1
    return s


def test_numpy_round_trip_is_byte_exact():
    s = small_snapshot()
    buf = s.to_bytes()
    t = Snapshot.from_bytes(buf)
    assert t.to_bytes() == buf
    assert t.sub_steps == 8 and t.polygon_contact and t.n_polygons == 2
    assert np.array_equal(bits(t.particles_inv_mass), bits(s.particles_inv_mass))
    g = t.polygon(1)
    assert len(g["pos"]) == 4 and len(g["link_ab"]) == 2 and not g["is_static"]
    assert np.array_equal(t.last_update, s.last_update)
    # an empty solver is a valid snapshot too
    assert Snapshot.from_bytes(Snapshot().to_bytes()).to_bytes() == Snapshot().to_bytes()


def test_numpy_reader_rejects_corrupt_files():
    buf = small_snapshot().to_bytes()
    with pytest.raises(SnapshotError):
        Snapshot.from_bytes(b"NOTASNAP" + buf[8:])
    with pytest.raises(SnapshotError):
        Snapshot.from_bytes(buf[:-4])
    with pytest.raises(SnapshotError):
        Snapshot.from_bytes(buf + b"\0\0\0\0")
    bad = small_snapshot()
    bad.particle_links_ab[0] = [5, 5]  # a < b (link.rs:19-21)
    with pytest.raises(SnapshotError):
        bad.to_bytes()
    bad = small_snapshot()
    bad.poly_links_ab[4] = [2, 4]  # polygon 1 has 4 points
    with pytest.raises(SnapshotError):
        bad.to_bytes()


def test_c_loader_validates_before_touching_a_device(tmp_path):
    """No GPU needed: a bad file is rejected by the parser; a good one gets as far as bendy_create."""
    L = _lib.lib()
    good = small_snapshot().to_bytes()
    cases = {
        "missing": None,
        "magic": b"NOTASNAP" + good[8:],
        "truncated": good[:-8],
        "trailing": good + b"\0" * 4,
    }
    bad_link = small_snapshot()
    raw = bytearray(good)
    # first particle link (0,1) -> (1,1): find it right after the particle block
    off = 144 + 50 * 8 * 2 + 50 * 4 + 3 * 8 * 3 + 3 * 4
    assert np.frombuffer(bytes(raw[off:off + 8]), u32).tolist() == [0, 1]
    raw[off:off + 4] = np.array([1], u32).tobytes()
    cases["link"] = bytes(raw)
    expect = {"missing": "cannot open", "magic": "bad magic", "truncated": "truncated", "trailing": "trailing",
              "link": "particle link out of range"}
    for name, data in cases.items():
        path = tmp_path / f"{name}.snap"
        if data is not None:
            path.write_bytes(data)
        h = L.bendy_load_snapshot(os.fsencode(str(path)), -1)
        assert not h, name
        assert expect[name] in L.bendy_last_error(None).decode(), name
    # a valid file passes the parser; without a GPU the failure is bendy_create's, loudly
    path = tmp_path / "good.snap"
    path.write_bytes(good)
    h = L.bendy_load_snapshot(os.fsencode(str(path)), -1)
    if h:
        L.bendy_destroy(h)
    else:
        assert "no CUDA device" in L.bendy_last_error(None).decode()


def test_oracle_replay_of_a_snapshot_equals_the_direct_load(tmp_path):
    """CPU only: Scene -> Snapshot -> file -> Snapshot -> oracle steps exactly like Scene -> oracle."""
    sc = scenes.c3_softbody_field(2, 2, 2, 3)
    sc.particles = (sc.particles - np.array([40.0, 0.0], f32)).astype(f32)
    path = tmp_path / "scene.snap"
    snapshot_from_scene(sc).save(str(path))
    snap = Snapshot.load(str(path))
    a = oracle_from_scene(sc)
    b = oracle_from_snapshot(snap, gravity=sc.gravity, bounds=sc.bounds)
    for _ in range(12):
        a.update(sc.dt)
        b.update(sc.dt)
    for x, y in zip(a.particles() + a.circles(), b.particles() + b.circles()):
        assert np.array_equal(bits(x), bits(y))
    for k in range(a.polygon_len()):
        for x, y in zip(a.polygon(k), b.polygon(k)):
            assert np.array_equal(bits(x), bits(y))


@pytest.mark.gpu
def test_saved_solver_continues_bit_identically(tmp_path):
    from bendy2d_b200 import Solver

    sc = scenes.c3_softbody_field(4, 3, 3, 4)
    g = Solver()
    sc.load_into(g)
    g.set_particle_inv_mass(np.linspace(0.5, 1.5, sc.n_particles).astype(f32))
    g.update(sc.dt, 5)
    path = str(tmp_path / "c3.snap")
    g.save_snapshot(path)
    h = Solver.load_snapshot(path)
    assert np.array_equal(h.gravity, g.gravity) and np.array_equal(h.bounds.size, g.bounds.size)
    # the numpy reader sees exactly the solver's state
    snap = Snapshot.load(path)
    pos, prev = g.read_particles()
    assert np.array_equal(bits(snap.particles_pos), bits(pos)) and np.array_equal(bits(snap.particles_prev), bits(prev))
    cp, cq, cr = g.read_circles()
    assert np.array_equal(bits(snap.circles_pos), bits(cp)) and np.array_equal(bits(snap.circles_radius), bits(cr))
    assert snap.n_polygons == g.get_polygons_len() and snap.sub_steps == sc.sub_steps
    assert snap.particles_inv_mass is not None and len(snap.particle_links_len) == sc.n_links
    # both solvers keep stepping: same bits
    g.update(sc.dt, 4)
    h.update(sc.dt, 4)
    for a, b in zip(g.read_particles() + g.read_circles(), h.read_particles() + h.read_circles()):
        assert np.array_equal(bits(a), bits(b))
    for k in range(g.get_polygons_len()):
        for a, b in zip(g.read_polygon(k)[:3], h.read_polygon(k)[:3]):
            assert np.array_equal(bits(a), bits(b))
    # a third copy rebuilt from the numpy reader through the public add_* calls agrees as well
    r = Solver()
    r.gravity, r.bounds = h.gravity.copy(), h.bounds
    snap.load_into(r)
    r.update(sc.dt, 4)
    for a, b in zip(g.read_particles(), r.read_particles()):
        assert np.array_equal(bits(a), bits(b))
    # and the file the clone writes is the file the numpy writer produces
    g2 = Solver.load_snapshot(path)
    p2 = str(tmp_path / "again.snap")
    g2.save_snapshot(p2)
    assert open(p2, "rb").read() == snap.to_bytes()


@pytest.mark.gpu
def test_snapshot_replays_into_the_oracle(tmp_path):
    """The use the format exists for: carry a GPU state to the CPU oracle and continue both."""
    from bendy2d_b200 import Solver

    sc = scenes.c3_softbody_field(3, 2, 2, 3)
    g = Solver()
    sc.load_into(g)
    g.update(sc.dt, 3)
    path = str(tmp_path / "state.snap")
    g.save_snapshot(path)
    snap = Snapshot.load(path)
    o = oracle_from_snapshot(snap)
    sync_schedule(g, o, sc)
    for _ in range(2):
        g.update(sc.dt)
        o.update(sc.dt)
    st = compare_state(g, o, scale=512.0, what="snapshot replay")
    assert st["ulp_pos"] == 0 and st["ulp_prev"] == 0, st


def test_corrupt_counts_are_refused_before_anything_is_allocated(tmp_path):
    """a header that announces 2^31 particles in a 200-byte file must give 'truncated snapshot', not an
    attempt to allocate 17 GB (std::bad_alloc through the C boundary)."""
    import struct

    s = snapshot_from_scene(scenes.c3_softbody_field(1, 1, 0, 0))
    raw = bytearray(s.to_bytes())
    off = 8 + 8 + 8 + 16 + 32  # magic, version/sub_steps, radius/cell, flags.., last[8] -> counts
    raw[off:off + 8] = struct.pack("<Q", 0x7FFFFFF0)
    p = tmp_path / "huge.b2d"
    p.write_bytes(bytes(raw))
    L = _lib.lib()
    h = L.bendy_load_snapshot(str(p).encode(), -1)
    assert not h
    msg = (L.bendy_last_error(None) or b"").decode()
    assert "truncated" in msg, msg
    raw = bytearray(s.to_bytes())
    raw[16:20] = struct.pack("<f", -1.0)  # particle_radius < 0: the setter would refuse it
    p.write_bytes(bytes(raw))
    assert not L.bendy_load_snapshot(str(p).encode(), -1)
    assert "particle radius" in (L.bendy_last_error(None) or b"").decode()
