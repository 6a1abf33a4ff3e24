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


def _inv_masses(n):
    rng = np.random.default_rng(5)
    k = rng.choice([0.25, 0.5, 1.0, 2.0, 4.0], n).astype(f32)
    k[rng.choice(n, 40, replace=False)] = 0.0  # pinned points
    return k


def _nccl_worker(rank, world, port, q, with_k=False):
    try:
        _nccl_worker_body(rank, world, port, q, with_k)
    except Exception as e:  # report instead of leaving the parent waiting for the queue
        q.put((rank, None, None, None, f"{type(e).__name__}: {e}"))
        raise


def _nccl_worker_body(rank, world, port, q, with_k):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sc = touching_field(24, 2)  # wide enough that every strip has interior bodies (overlapped exchange path)
    sv = strips.StripSolver(sc, rank, world, rank, dist)
    if with_k:
        sv.set_particle_inv_mass(_inv_masses(sc.n_particles))
    info = sv.schedule_info()
    sv.update(sc.dt, n=60)
    pos, prev = sv.read_particles()
    sv.check_halo()
    q.put((rank, sv.part.global_index, pos, prev, info))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("with_k", [False, True])
def test_nccl_strips_match_single_gpu(with_k):
    """with_k: ext inverse masses - the packed discs' scales travel with their positions (two more messages)"""
    import torch

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    sc = touching_field(24, 2)
    ref = Solver(0)
    sc.load_into(ref)
    if with_k:
        ref.set_particle_inv_mass(_inv_masses(sc.n_particles))
    ref.update(sc.dt, n=60)
    rp, rq = ref.read_particles()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (3 if with_k else 0)
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, q, with_k)) for r in range(world)]
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


def _nccl_cut_worker(rank, world, port, reference_order, q):
    import torch
    import torch.distributed as dist

    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from test_emu_strips_mp import _cut_bodies_case

        ok, info = _cut_bodies_case(rank, world, dist, reference_order, device=rank)
        q.put((rank, ok, info))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:
        q.put((rank, False, f"{type(e).__name__}: {e}"))
        raise


@pytest.mark.parametrize("reference_order", [False, True])
def test_nccl_bodies_cut_by_strip_edges(reference_order):
    """N6 on real NCCL: links across strip edges; sharded = oracle (replayed order) and, in reference order,
    = the unsharded GPU run bit for bit"""
    import torch

    world = min(torch.cuda.device_count(), 3)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (7 if reference_order else 0)
    procs = [ctx.Process(target=_nccl_cut_worker, args=(r, world, port, reference_order, q)) for r in range(world)]
    for p in procs:
        p.start()
    for _ in procs:
        rank, ok, info = q.get(timeout=300)
        assert ok is True, f"rank {rank}: {info}"
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0


def _nccl_rebalance_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        # free particles pile up and spread sideways: ownership has to follow (rebalance over the process group);
        # replicated Circles in the pile: their state must survive the re-partition
        sc = scenes.c2_free_particles(80, 30)
        sc.bounds = (0.0, 0.0, 48.0, 12.0)
        sc.circles_pos = np.array([[18.0, 10.5], [24.0, 10.8], [30.0, 10.2]], f32)
        sc.circles_r = np.array([0.9, 0.6, 1.1], f32)
        sv = strips.StripSolver(sc, rank, world, rank, dist, band=4.0)
        n_updates, rebalanced = 150, 0
        for _ in range(n_updates):
            sv.update(sc.dt)
            if sv.needs_rebalance():
                sv.rebalance()
                rebalanced += 1
        pos, prev = sv.read_particles()
        gpos, gprev = strips.gather_global_state(dist, world, rank, sv.part.global_index, pos, prev, sc.n_particles)
        ok = True
        if rank == 0:
            ref = Solver(rank)
            sc.load_into(ref)
            ref.update(sc.dt, n=n_updates)
            rp, rq = ref.read_particles()
            ok = max_ulp(gpos, rp) == 0 and max_ulp(gprev, rq) == 0
            a, b = sv.read_circles(), ref.read_circles()
            ok = ok and max_ulp(a[0], b[0]) == 0 and max_ulp(a[1], b[1]) == 0
        q.put((rank, ok, rebalanced))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:
        q.put((rank, False, f"{type(e).__name__}: {e}"))
        raise


def test_nccl_rebalance_over_the_process_group():
    """StripSolver.rebalance() on real NCCL (only the CPU emulation had run it): re-partition by the
    current positions, bit-identical continuation, replicated Circles carried across"""
    import torch

    world = min(torch.cuda.device_count(), 3)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_nccl_rebalance_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    for _ in procs:
        res.append(q.get(timeout=600))
        assert res[-1][1] is True, f"rank {res[-1][0]}: {res[-1][2]}"
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert max(r[2] for r in res) > 0, "the scene never needed rebalancing: nothing was tested"
