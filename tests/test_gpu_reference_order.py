"""Reference-order link schedule (bendy_set_link_schedule(BENDY_LINKS_REFERENCE_ORDER)).

The default schedule colours the links greedily: its results equal the reference fed the links in the exported
colour order (the north_star contract).  In reference-order mode the colours are dependency levels in insertion
order, so the device must reproduce the reference's OWN walk (solver.rs:143-146: `for link in particle_links`,
insertion order) bit for bit - the oracle below is NOT given the device's link order."""
import os

import numpy as np
import pytest

from bendy2d_b200 import Solver, scenes
from helpers import bits, compare_state, f32, max_ulp, oracle_from_scene

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(sc, **kw):
    g = Solver()
    g.set_link_schedule("reference")
    for k, v in kw.items():
        getattr(g, k)(*v)
    sc.load_into(g)
    return g


def test_c1_in_reference_order_reproduces_the_insertion_order_golden_bit_for_bit():
    gold = np.load(os.path.join(G, "c1_reference_order.npz"))
    sc = scenes.c1_softbody_blob()
    g = _load(sc)
    g.update(sc.dt, n=int(gold["n_updates"]))
    pos, prev = g.read_particles()
    cp, cq, _ = g.read_circles()
    assert np.array_equal(bits(pos), bits(gold["pos"])) and np.array_equal(bits(prev), bits(gold["prev"]))
    assert np.array_equal(bits(cp), bits(gold["circle_pos"]))
    info = g.schedule_info()
    assert info["n_global_links"] == 0 and info["n_local_colours"] > 8  # levels, not greedy colours


def test_c1_multi_kernel_path_in_reference_order(monkeypatch):
    monkeypatch.setenv("BENDY_SMALL_SCENE", "0")
    gold = np.load(os.path.join(G, "c1_reference_order.npz"))
    sc = scenes.c1_softbody_blob()
    g = _load(sc)
    g.update(sc.dt, n=int(gold["n_updates"]))
    assert np.array_equal(bits(g.read_particles()[0]), bits(gold["pos"]))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_shuffled_link_lists_match_the_insertion_order_oracle(seed):
    """Bodies whose links were added in a random order: the dependency levels follow THAT order."""
    rng = np.random.default_rng(seed)
    sc = scenes.c3_softbody_field(3, 2, 0, 0)
    sc.particle_radius = 0.0  # reference semantics: free particles collide with nothing
    perm = rng.permutation(sc.n_links)
    sc.links_ab, sc.links_len = sc.links_ab[perm].copy(), sc.links_len[perm].copy()
    g = _load(sc)
    o = oracle_from_scene(sc)  # insertion order, no schedule replay
    for k in range(12):
        g.update(sc.dt)
        o.update(sc.dt)
        st = compare_state(g, o, 128.0, 1e-5, f"shuffled links update {k + 1}")
        assert st["ulp_pos"] == 0 and st["ulp_prev"] == 0, st


def test_links_crossing_partitions_fall_back_to_one_launch_per_level():
    """A body larger than a partition: the levels are taken over the whole graph (all links 'global')."""
    sc = scenes.c3_softbody_field(2, 1, 0, 0)
    sc.particle_radius = 0.0
    g = _load(sc, set_plan_params=(64, 128))
    o = oracle_from_scene(sc)
    info = g.schedule_info()
    assert info["n_local_links"] == 0 and info["n_global_links"] == sc.n_links and info["n_global_colours"] > 8
    for k in range(6):
        g.update(sc.dt)
        o.update(sc.dt)
    assert max_ulp(g.read_particles()[0], o.particles()[0]) == 0


def test_a_long_chain_needs_one_level_per_link():
    """300 links added end to end: 300 dependency levels (> the partition kernel's 255 colours)."""
    n = 301
    pos = np.stack([np.linspace(10, 70, n), np.full(n, 20.0)], 1).astype(f32)
    ab = np.stack([np.arange(n - 1), np.arange(1, n)], 1).astype(np.uint32)
    ln = np.full(n - 1, 0.15, f32)
    sc = scenes.Scene(name="chain", particles=pos, links_ab=ab, links_len=ln, bounds=(0.0, 0.0, 100.0, 100.0))
    g = _load(sc)
    o = oracle_from_scene(sc)
    assert g.schedule_info()["n_global_colours"] == n - 1
    for k in range(5):
        g.update(sc.dt)
        o.update(sc.dt)
    assert max_ulp(g.read_particles()[0], o.particles()[0]) == 0


def test_disc_contacts_on_top_of_reference_order_links():
    sc = scenes.c3_softbody_field(3, 2, 2, 2)
    g = _load(sc)
    o = oracle_from_scene(sc)
    o.set_point_rank(g.point_rank())
    o.set_grid(*g.grid())
    for k in range(20):
        g.update(sc.dt)
        o.update(sc.dt)
    st = compare_state(g, o, 128.0, 1e-5, "reference order + discs")
    assert st["ulp_pos"] == 0


def test_clone_and_snapshot_keep_the_schedule(tmp_path):
    sc = scenes.c1_softbody_blob()
    g = _load(sc)
    g.update(sc.dt, n=3)
    c = g.clone()
    path = str(tmp_path / "ref.b2d")
    g.save_snapshot(path)
    r = Solver.load_snapshot(path)
    for s_ in (g, c, r):
        s_.update(sc.dt, n=5)
    a = bits(g.read_particles()[0])
    assert np.array_equal(a, bits(c.read_particles()[0])) and np.array_equal(a, bits(r.read_particles()[0]))
    assert r.schedule_info()["n_local_colours"] == g.schedule_info()["n_local_colours"]
