"""CPU tests of the strip sharding logic (no GPU): partition invariants, and a 2-process gloo run
that moves the halo bands between ranks and checks that every contact partner is visible."""
import os
import sys

import numpy as np
import pytest

from bendy2d_b200 import scenes, strips

f32 = np.float32


def small_field():
    sc = scenes.c3_softbody_field(8, 3, 0, 0)
    # squeeze the columns together so that neighbouring bodies really touch across strip edges
    col = (np.arange(sc.n_particles) // 500) % 8
    sc.particles[:, 0] -= (col * 3.1).astype(f32)
    return sc


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_invariants(world):
    sc = small_field()
    parts = strips.partition_scene(sc, world)
    owned = np.concatenate([p.global_index for p in parts])
    assert sorted(owned.tolist()) == list(range(sc.n_particles)), "every particle owned exactly once"
    assert sum(p.scene.n_links for p in parts) == sc.n_links
    for p in parts:
        ab = p.scene.links_ab.astype(np.int64)
        assert (ab[:, 0] < ab[:, 1]).all() and ab.max() < p.scene.n_particles
        g = p.global_index[ab]
        d = sc.particles[g[:, 0]] - sc.particles[g[:, 1]]
        np.testing.assert_allclose(np.hypot(d[:, 0], d[:, 1]), p.scene.links_len, rtol=1e-6)
        assert p.x_left < p.x_right
        if world > 1:
            assert p.ghost_cap == parts[0].ghost_cap > 0
    for a, b in zip(parts[:-1], parts[1:]):
        assert a.x_right == b.x_left
    assert parts[0].x_left == -np.inf and parts[-1].x_right == np.inf
    # the automatic body detection agrees with the generator's body ids
    assert len(np.unique(strips.body_ids(sc))) == len(np.unique(sc.body_of))


def test_partition_replicates_circles_and_polygons():
    sc = scenes.c3_softbody_field(4, 1, 3, 5)
    parts = strips.partition_scene(sc, 2)
    for p in parts:
        assert len(p.scene.polygons) == 5 and p.scene.polygons_static == sc.polygons_static
        assert p.scene.polygon_contact == sc.polygon_contact
        assert np.array_equal(p.scene.circles_pos, sc.circles_pos) and np.array_equal(p.scene.circles_r, sc.circles_r)


def _halo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = small_field()
    part = strips.partition_scene(sc, world)[rank]
    px = part.scene.particles
    cap = part.ghost_cap
    ghosts = []
    reqs, bufs = [], []
    for side, peer, mask in ((0, rank - 1, px[:, 0] < part.send_left_below),
                             (1, rank + 1, px[:, 0] > part.send_right_above)):
        if not 0 <= peer < world:
            continue
        send = torch.full((cap, 2), float("nan"))
        sel = torch.from_numpy(px[mask])
        assert len(sel) <= cap
        send[: len(sel)] = sel
        recv = torch.empty((cap, 2))
        reqs += [dist.isend(send, peer), dist.irecv(recv, peer)]
        bufs.append(recv)
    for r in reqs:
        r.wait()
    for b in bufs:
        g = b.numpy()
        ghosts.append(g[np.isfinite(g[:, 0])])
    local = np.concatenate([px] + ghosts) if ghosts else px
    # every particle of the full scene within contact range of an owned particle must be local
    r2 = (2 * sc.particle_radius + 1e-4) ** 2
    full = sc.particles
    missing = 0
    for i in range(0, len(px), 997):  # sample owned particles
        d2 = ((full - px[i]) ** 2).sum(1)
        for j in np.nonzero(d2 < r2)[0]:
            if not (np.abs(local - full[j]).sum(1) == 0).any():
                missing += 1
    q.put((rank, missing, sum(len(g) for g in ghosts)))
    dist.destroy_process_group()


def test_halo_bands_cover_all_contacts_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, missing, n_ghost in res:
        assert missing == 0, f"rank {rank}: {missing} contact partners not covered by the halo"
        assert n_ghost > 0


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 10007
    rng = np.random.default_rng(11)  # same stream on every rank
    pos = rng.uniform(-50, 50, (n, 2)).astype(f32)
    prev = rng.uniform(-50, 50, (n, 2)).astype(f32)
    pos[5] = [np.nan, -0.0]
    prev[9] = [np.inf, 1e-42]  # a denormal must survive too
    owner = rng.integers(0, world, n)
    owner[:100] = 0  # uneven shares
    mine = np.nonzero(owner == rank)[0]
    mine = mine[rng.permutation(len(mine))]  # the local (internal) order is not the user order
    gpos, gprev = strips.gather_global_state(dist, world, 0, mine, pos[mine], prev[mine], n)
    ok = np.array_equal(gpos.view(np.uint32), pos.view(np.uint32)) and np.array_equal(gprev.view(np.uint32), prev.view(np.uint32))
    # a particle nobody owns / two owners must be reported, not silently filled with garbage
    bad = mine[1:] if rank == 0 else mine
    try:
        strips.gather_global_state(dist, world, 0, bad, pos[bad], prev[bad], n)
        caught = False
    except strips.HaloError:
        caught = True
    q.put((rank, ok, caught))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rebalance_gather_is_bit_exact_gloo(world):
    """StripSolver.rebalance() re-partitions from the gathered global state: the gather (the part that needs
    the process group) on CPU tensors under gloo, uneven shares, shuffled local order, NaN / -0 / denormals"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and caught for _, ok, caught in res), res


def test_stray_margins_shrink_next_to_a_narrow_strip_and_too_narrow_strips_are_refused():
    """The exchange only reaches the two neighbours, so the discs of strips k-1 and k+1 must never be able to
    touch inside strip k: next to a narrow strip the stray margins shrink to half of what it leaves, and a
    strip that cannot keep them 2r apart is refused (found by the randomised strip fields: two bodies owned
    by strips 1 and 3 touched inside a 1.6-wide strip 2 and neither owner saw the other)."""
    sc = small_field()  # columns of bodies 4.9 apart, default band 6.95
    parts = strips.partition_scene(sc, 8)
    r2 = 2 * sc.particle_radius
    for k, p in enumerate(parts):
        assert (p.stray_margin_left is None) == (k == 0) and (p.stray_margin_right is None) == (k == 7)
    for a, b, c in zip(parts, parts[1:], parts[2:]):  # a strays right into b, c strays left into b
        width = b.x_right - b.x_left
        assert a.stray_margin_right == c.stray_margin_left == min(0.5 * b.band, 0.5 * (width - r2))
        assert (a.stray_right - b.x_left) + (b.x_right - c.stray_left) + r2 <= width + 1e-9
    # wide strips keep band / 2 (the 16M benchmark scene: strips 256 wide, band 6.95)
    wide = strips.partition_scene(sc, 2)
    assert wide[0].stray_right == wide[0].x_right + 0.5 * wide[0].band
    # a strip narrower than the contact range cannot be protected at all
    sc2 = scenes.c2_free_particles(40, 4)
    sc2.particles[:, 0] = (10.0 + 0.001 * np.arange(sc2.n_particles)).astype(f32)  # 160 discs within 0.16
    with pytest.raises(ValueError, match="too narrow"):
        strips.partition_scene(sc2, 4, band=1.0)
