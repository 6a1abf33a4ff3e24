"""GPU tests of the multi-GPU strip path.

1-GPU: LocalStripGroup (all strips in one process, device-to-device halo copies, the same kernels
and arithmetic as the NCCL path) must reproduce the unsharded run bit for bit.
>=2 GPUs: the NCCL path, launched with torch.multiprocessing, against the unsharded run.
"""
import os

import numpy as np
import pytest

from bendy2d_b200 import Solver, scenes, strips
from helpers import bits, max_ulp

pytestmark = pytest.mark.gpu
f32 = np.float32


def touching_field(nx=8, ny=3):
    sc = scenes.c3_softbody_field(nx, ny, 0, 0)
    col = (np.arange(sc.n_particles) // 500) % nx
    sc.particles[:, 0] -= (col * 3.1).astype(f32)  # neighbouring bodies overlap slightly -> contacts at once
    sc.bounds = (0.0, 0.0, 128.0, 48.0)
    return sc


@pytest.mark.parametrize("n_strips", [2, 3, 5])
def test_local_strip_group_is_bit_identical_to_single_solver(n_strips):
    # 5 strips need a wider field: next to a strip that holds a single column of bodies the stray margin
    # shrinks below the half width of a body (strips.partition_scene) and the run would be flagged at once
    sc = touching_field() if n_strips < 5 else touching_field(16, 2)
    ref = Solver()
    sc.load_into(ref)
    grp = strips.LocalStripGroup(sc, n_strips)
    for k in range(6):
        ref.update(sc.dt, n=10)
        grp.update(sc.dt, n=10)
        rp, rq = ref.read_particles()
        gp, gq = grp.read_particles()
        assert max_ulp(gp, rp) == 0 and max_ulp(gq, rq) == 0, f"after {10 * (k + 1)} substeps"
    stats = grp.halo_stats()
    assert all(o == 0 and st == 0 for _, _, o, st in stats), stats
    assert sum(a + b for a, b, _, _ in stats) > 0, "no halo traffic: the test scene does not exercise the exchange"
    # contacts across strip edges really happened: positions differ from a run without collisions
    free = Solver()
    sc2 = touching_field() if n_strips < 5 else touching_field(16, 2)
    sc2.particle_radius = 0.0
    sc2.load_into(free)
    free.update(sc.dt, n=60)
    assert not np.array_equal(bits(free.read_particles()[0]), bits(rp))


def test_strip_group_free_particles_with_rebalancing():
    """Free particles pile up and spread sideways, so ownership has to follow the positions: the
    stray flag is polled every update (max speed ~45 units/s = 0.4 per substep << band/2)."""
    sc = scenes.c2_free_particles(80, 30)
    sc.bounds = (0.0, 0.0, 48.0, 12.0)
    ref = Solver()
    sc.load_into(ref)
    grp = strips.LocalStripGroup(sc, 4, band=4.0)
    rebalanced = 0
    for _ in range(150):
        ref.update(sc.dt)
        grp.update(sc.dt)
        if grp.needs_rebalance():
            grp.rebalance()
            rebalanced += 1
    assert all(o == 0 for _, _, o, _ in grp.halo_stats())
    gp, gq = grp.read_particles()
    rp, rq = ref.read_particles()
    assert max_ulp(gp, rp) == 0 and max_ulp(gq, rq) == 0, f"rebalanced {rebalanced} times"


def test_strip_group_sub_steps():
    sc = touching_field(6, 2)
    sc.sub_steps, sc.dt = 4, 4.0 / 120.0
    ref = Solver()
    sc.load_into(ref)
    grp = strips.LocalStripGroup(sc, 3)
    ref.update(sc.dt, n=10)
    grp.update(sc.dt, n=10)
    assert max_ulp(grp.read_particles()[0], ref.read_particles()[0]) == 0


def _nccl_worker(rank, world, port, q):
    try:
        _nccl_worker_body(rank, world, port, q)
    except Exception as e:  # report instead of leaving the parent waiting for the queue
        q.put((rank, None, None, None, f"{type(e).__name__}: {e}"))
        raise


def _nccl_worker_body(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sc = touching_field(24, 2)  # wide enough that every strip has interior bodies (overlapped exchange path)
    sv = strips.StripSolver(sc, rank, world, rank, dist)
    info = sv.schedule_info()
    sv.update(sc.dt, n=60)
    pos, prev = sv.read_particles()
    sv.check_halo()
    q.put((rank, sv.part.global_index, pos, prev, info))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_strips_match_single_gpu():
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    sc = touching_field(24, 2)
    ref = Solver(0)
    sc.load_into(ref)
    ref.update(sc.dt, n=60)
    rp, rq = ref.read_particles()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gp, gq = np.empty_like(rp), np.empty_like(rq)
    for _ in procs:
        rank, idx, pos, prev, info = q.get(timeout=300)
        assert idx is not None, f"rank {rank} failed: {info}"
        gp[idx], gq[idx] = pos, prev
        assert info["n_partitions"] > 0, info
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert max_ulp(gp, rp) == 0 and max_ulp(gq, rq) == 0


def test_strips_with_replicated_polygons_match_single_solver():
    """Polygons are replicated on every strip (nothing a particle does reaches a polygon): bodies rain on a row of
    static obstacles and on a dynamic overlapping polygon pair across 3 strips; the particles AND every strip's
    copy of the polygons must equal the unsharded run bit for bit."""
    sc = scenes.c3_softbody_field(8, 2, 0, 12)
    col = (np.arange(sc.n_particles) // 500) % 8
    sc.particles[:, 0] -= (col * 3.1).astype(f32)
    sc.bounds = (0.0, 0.0, 128.0, 64.0)
    # the obstacles right under the bodies (x 56..95, y 8..23); two of them dynamic and overlapping
    placed = []
    for j, p in enumerate(sc.polygons):
        placed.append((p - p.mean(0) + np.array([57.0 + 3.3 * j, 27.5])).astype(f32))
    placed[1] = (placed[0] + np.array([1.0, 0.5], f32)).astype(f32)
    sc.polygons = placed
    sc.polygons_static = [False, False] + [True] * (len(placed) - 2)
    ref = Solver()
    sc.load_into(ref)
    grp = strips.LocalStripGroup(sc, 3)
    checked = 0
    for k in range(5):
        ref.update(sc.dt, n=20)
        grp.update(sc.dt, n=20)
        if any(o or st for _, _, o, st in grp.halo_stats()):
            break  # a body bounced further sideways than the stray margin: from here on the run needs a rebalance
        rp, rq = ref.read_particles()
        gp, gq = grp.read_particles()
        assert max_ulp(gp, rp) == 0 and max_ulp(gq, rq) == 0, f"after {20 * (k + 1)} substeps"
        for sv in grp.solvers:
            for j in range(len(sc.polygons)):
                a, b = sv.read_polygon(j), ref.read_polygon(j)
                assert max_ulp(a[0], b[0]) == 0 and max_ulp(a[1], b[1]) == 0 and max_ulp(a[2], b[2]) == 0, (k, j)
        checked = k + 1
    assert checked >= 2, "the halo went stale before the bodies reached the obstacles"
    free = Solver()  # the obstacles really were hit: without them the particles end up elsewhere
    sc2 = scenes.Scene(sc.name, sc.bounds, particle_radius=sc.particle_radius, particles=sc.particles,
                       links_ab=sc.links_ab, links_len=sc.links_len)
    sc2.load_into(free)
    free.update(sc.dt, n=20 * checked)
    assert not np.array_equal(bits(free.read_particles()[0]), bits(rp))


def test_strips_with_replicated_circles_match_single_solver():
    """Circles are replicated; the fixed-point corrections each strip's own discs collect for a Circle are summed
    over the strips before the Circles' tail applies them (integer sums: same bits as the unsharded sum).  Bodies
    fall onto a row of Circles (two of them linked, two overlapping) across 3 strips."""
    sc = scenes.c3_softbody_field(8, 2, 0, 0)
    col = (np.arange(sc.n_particles) // 500) % 8
    sc.particles[:, 0] -= (col * 3.1).astype(f32)
    sc.bounds = (0.0, 0.0, 128.0, 64.0)
    sc.circles_pos = np.stack([57.0 + 3.9 * np.arange(10), np.full(10, 26.0)], 1).astype(f32)
    sc.circles_pos[3] = sc.circles_pos[2] + np.array([0.8, 0.3], f32)  # an overlapping pair: the exact circle pass runs
    sc.circles_r = np.array([1.2, 0.9, 1.1, 1.0, 1.4, 0.8, 1.3, 1.0, 0.7, 1.2], f32)
    ref = Solver()
    sc.load_into(ref)
    grp = strips.LocalStripGroup(sc, 3)
    checked = 0
    for k in range(5):
        ref.update(sc.dt, n=20)
        grp.update(sc.dt, n=20)
        if any(o or st for _, _, o, st in grp.halo_stats()):
            break
        rp, rq = ref.read_particles()
        gp, gq = grp.read_particles()
        assert max_ulp(gp, rp) == 0 and max_ulp(gq, rq) == 0, f"after {20 * (k + 1)} substeps"
        rc = ref.read_circles()
        for sv in grp.solvers:
            gc = sv.read_circles()
            assert max_ulp(gc[0], rc[0]) == 0 and max_ulp(gc[1], rc[1]) == 0, k
        checked = k + 1
    assert checked >= 3, "the halo went stale before the bodies had pushed the circles around"
    still = Solver()  # the particles really pushed the Circles: alone they would be somewhere else
    still.add_circles(sc.circles_pos, sc.circles_r)
    still.bounds.size[:] = (128.0, 64.0)
    still.update(sc.dt, n=20 * checked)
    assert not np.array_equal(bits(still.read_circles()[0]), bits(rc[0]))
