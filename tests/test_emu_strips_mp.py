"""The multi-process strip path (bendy2d_b200.strips.StripSolver: one process per GPU, NCCL halo exchange issued
inside the captured graph, rebalancing over the process group) WITHOUT GPUs: every rank process loads the CPU
emulation build of the library (tests/cuemu) and a stub of the eight NCCL entry points that moves the messages
over FIFOs (tests/cuemu/nccl_stub.cpp); torch.distributed runs on gloo.  Test infrastructure only; what it
checks is the host logic of the N>1 path (partition, unique-id broadcast, communicator, send/recv inside the
graph, stray detection, rebalance) and that the sharded run is bit-identical to the unsharded one.
"""
import os
import shutil
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "cuemu"))

pytestmark = pytest.mark.skipif(shutil.which(os.environ.get("CXX", "g++")) is None, reason="needs g++")
f32 = np.float32


def _worker(rank, world, port, case, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, HERE)
    from bendy2d_b200 import Solver, _lib, scenes, strips
    from test_gpu_strips import touching_field

    assert _lib.LIB_PATH.endswith("_emu.so")
    try:
        if case in ("cut", "cut_ref"):
            ok, info = _cut_bodies_case(rank, world, dist, case == "cut_ref")
            q.put((rank, ok, info, 0, 1))
            return
        kmass = None
        if case in ("bodies", "polygons", "circles", "kmass"):
            sc = touching_field() if world < 4 else touching_field(16, 2)
            if case == "kmass":  # ext inverse masses: the packed discs' scales travel with their positions
                rng = np.random.default_rng(5)
                kmass = rng.choice([0.25, 0.5, 1.0, 2.0, 4.0], sc.n_particles).astype(f32)
                kmass[rng.choice(sc.n_particles, 40, replace=False)] = 0.0
            if case == "polygons":
                sc = touching_field(8, 2)  # two rows of bodies (y 8..23) above the obstacles  # replicated obstacles under the bodies, one dynamic overlapping pair
                src = scenes.c3_softbody_field(2, 1, 0, 10)
                placed = [(p - p.mean(0) + np.array([57.0 + 3.3 * j, 27.5])).astype(f32) for j, p in enumerate(src.polygons)]
                placed[1] = (placed[0] + np.array([1.0, 0.5], f32)).astype(f32)
                sc.polygons, sc.polygons_static = placed, [False, False] + [True] * (len(placed) - 2)
                sc.polygon_contact = True
                sc.bounds = (0.0, 0.0, 128.0, 64.0)
            if case == "circles":  # replicated Circles under the bodies: their corrections are all-reduced
                sc = touching_field(8, 2)
                sc.bounds = (0.0, 0.0, 128.0, 64.0)
                sc.circles_pos = np.stack([57.0 + 3.9 * np.arange(10), np.full(10, 26.0)], 1).astype(f32)
                sc.circles_pos[3] = sc.circles_pos[2] + np.array([0.8, 0.3], f32)
                sc.circles_r = np.array([1.2, 0.9, 1.1, 1.0, 1.4, 0.8, 1.3, 1.0, 0.7, 1.2], f32)
            sv = strips.StripSolver(sc, rank, world, 0, dist)
            if kmass is not None:
                sv.set_particle_inv_mass(kmass)
            rounds = 3 if case == "bodies" else 2  # later the bodies bounce off the obstacles past the stray margin
            for _ in range(rounds):
                sv.update(sc.dt, n=20)
                sent = sv.check_halo()
            n_updates, rebalanced = 20 * rounds, 0
        else:  # free particles pile up and spread sideways: ownership has to follow (rebalance over the group)
            sc = scenes.c2_free_particles(80, 30)
            sc.bounds = (0.0, 0.0, 48.0, 12.0)
            if case == "free_circles":  # replicated Circles in the pile: their state must survive the re-partition
                sc.circles_pos = np.array([[18.0, 10.5], [24.0, 10.8], [30.0, 10.2]], f32)
                sc.circles_r = np.array([0.9, 0.6, 1.1], f32)
            sv = strips.StripSolver(sc, rank, world, 0, dist, band=4.0)
            n_updates, rebalanced, sent = 150, 0, (0, 0)
            for _ in range(n_updates):
                sv.update(sc.dt)
                if sv.needs_rebalance():
                    sv.rebalance()
                    rebalanced += 1
        pos, prev = sv.read_particles()
        gpos, gprev = strips.gather_global_state(dist, world, 0, sv.part.global_index, pos, prev, sc.n_particles)
        ok = True
        if rank == 0 or case in ("circles", "free_circles"):
            ref = Solver()
            sc.load_into(ref)
            if kmass is not None:
                ref.set_particle_inv_mass(kmass)
            ref.update(sc.dt, n=n_updates)
            rp, rq = ref.read_particles()
            ok = np.array_equal(gpos.view(np.uint32), rp.view(np.uint32)) and np.array_equal(gprev.view(np.uint32), rq.view(np.uint32))
            if case in ("circles", "free_circles"):  # every rank's copy of the Circles equals the unsharded ones
                a, b = sv.read_circles(), ref.read_circles()
                ok = ok and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
                moved = not np.array_equal(b[0][:, 0], sc.circles_pos[:, 0])  # the discs pushed them sideways
                ok = ok and moved
        q.put((rank, ok, int(sum(sent)), rebalanced, sv.schedule_info()["kernels_per_substep"]))
    except Exception as e:  # report instead of hanging the other ranks' queue reader
        q.put((rank, False, repr(e), 0, 0))
        raise
    finally:
        dist.destroy_process_group()


def wide_bodies_scene():
    """two lattice bodies 12.5 units wide (51 x 5 points, spacing 0.25) one above the other, touching discs: any cut
    into 2 or 3 strips of equal particle count goes through both bodies"""
    from bendy2d_b200 import scenes

    pos0, ab0 = scenes.lattice_body(51, 5, 0.25, (0.0, 0.0), True)
    pos = np.concatenate([pos0 + np.array([10.0, 6.0]), pos0 * np.array([1.0, 0.78]) + np.array([14.0, 7.3])]).astype(f32)
    ab = np.concatenate([ab0, ab0 + len(pos0)]).astype(np.uint32)
    d = pos[ab[:, 0]] - pos[ab[:, 1]]
    ln = np.hypot(d[:, 0], d[:, 1]).astype(f32)
    ln[len(ab0):] = (ln[len(ab0):] * f32(1.1)).astype(f32)  # the second body starts compressed: its links work from update 1
    return scenes.Scene(name="wide bodies", bounds=(0.0, 0.0, 40.0, 12.0), particles=pos, links_ab=ab, links_len=ln,
                        particle_radius=0.1)


def _cut_bodies_case(rank, world, dist, reference_order, device=0):
    """bodies cut by the strip edges: links across the edges run as trailing colours with an exchange of the endpoint
    positions before each.  coloured schedule: the oracle replays the sharded run's sequential order; reference
    order: with the scene's links listed in that order, sharded = unsharded = the oracle's insertion-order walk."""
    from bendy2d_b200 import Solver, strips
    from helpers import oracle_from_scene

    sc = wide_bodies_scene()
    n_updates = 40
    if reference_order:
        # list the links as [strip 0's][strip 1's] ... [cross, colour-major]: the order the sharded run executes
        parts = strips.partition_scene(sc, world, cut_bodies=True)
        order = strips.sequential_link_order(parts, [np.arange(len(p.local_links)) for p in parts])
        sc.links_ab, sc.links_len = sc.links_ab[order].copy(), sc.links_len[order].copy()
    sv = strips.StripSolver(sc, rank, world, device, dist, cut_bodies=True,
                            link_schedule="reference" if reference_order else "coloured")
    assert sv.part.cross is not None and len(sv.part.cross["mine"]) > 0, "the cut missed the bodies"
    sv.update(sc.dt, n=n_updates)
    sv.check_halo()
    pos, prev = sv.read_particles()
    gpos, gprev = strips.gather_global_state(dist, world, device, sv.part.global_index, pos, prev, sc.n_particles)
    orders = [None] * world
    dist.all_gather_object(orders, sv.link_order())
    if rank != 0:
        return True, "-"
    parts = strips.partition_scene(sc, world, cut_bodies=True)
    o = oracle_from_scene(sc)
    o.set_grid(*sv.grid())
    if not reference_order:
        o.set_link_order(strips.sequential_link_order(parts, orders))
    for _ in range(n_updates):
        o.update(sc.dt)
    op, oq = o.particles()
    ok = np.array_equal(gpos.view(np.uint32), op.view(np.uint32)) and np.array_equal(gprev.view(np.uint32), oq.view(np.uint32))
    info = f"cross links {sum(len(p.cross['mine']) for p in parts) // 2}, colours {parts[0].cross['n_colours']}"
    if reference_order:
        ref = Solver(device)
        ref.set_link_schedule("reference")
        sc.load_into(ref)
        ref.update(sc.dt, n=n_updates)
        rp, rq = ref.read_particles()
        ok = ok and np.array_equal(gpos.view(np.uint32), rp.view(np.uint32)) and np.array_equal(gprev.view(np.uint32), rq.view(np.uint32))
    moved = float(np.abs(gpos - sc.particles).max())
    return bool(ok and moved > 0.5), info + f", max displacement {moved:.2f}"


def _run(world, case, extra_env=None):
    import build as cuemu_build
    import torch.multiprocessing as mp

    lib = cuemu_build.build()
    env = {"BENDY2D_B200_LIB": lib, "BENDY_CUDA_EMU": "1", "BENDY_NCCL_LIB": cuemu_build.NCCL_STUB}
    env.update(extra_env or {})
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    procs = []
    try:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = 33500 + (os.getpid() % 2000) + 7 * world + {"bodies": 11, "polygons": 23, "circles": 37, "free_circles": 51, "cut": 63, "cut_ref": 77, "kmass": 91}.get(case, 0)
        procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = []
        for _ in procs:
            res.append(q.get(timeout=300))
            if res[-1][1] is not True:  # a rank failed: its peers are stuck in the exchange, do not wait for them
                for p in procs:
                    p.join(timeout=2)
                    if p.is_alive():
                        p.kill()
                raise AssertionError(f"rank {res[-1][0]} failed: {res[-1][2]}")
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        import glob

        for d in glob.glob("/tmp/cuemu_nccl_*"):  # the stub's FIFO directories (one per communicator)
            if any(d.startswith(f"/tmp/cuemu_nccl_{p.pid}_") for p in procs):
                shutil.rmtree(d, ignore_errors=True)
    return sorted(res)


@pytest.mark.parametrize("world,extra", [(2, {}), (3, {"BENDY_PDL_NCCL": "0"}), (4, {"BENDY_HALO_OVERLAP": "1"})])
def test_nccl_strip_solvers_match_the_single_solver_bit_for_bit(world, extra):
    res = _run(world, "bodies", extra)
    assert all(ok is True for _, ok, *_ in res), res
    assert sum(r[2] for r in res) > 0, "no halo traffic: the scene does not exercise the exchange"
    assert all(r[4] > 0 for r in res), "the substeps did not run as a captured graph"


@pytest.mark.parametrize("case", ["free", "free_circles"])
def test_nccl_strip_solvers_rebalance_over_the_process_group(case):
    res = _run(3, case)
    assert all(ok is True for _, ok, *_ in res), res
    assert res[0][3] > 0, "the scene never needed rebalancing: nothing was tested"


def test_nccl_strip_solvers_with_replicated_polygons():
    res = _run(2, "polygons")
    assert all(ok is True for _, ok, *_ in res), res


@pytest.mark.parametrize("world", [2, 3])
def test_nccl_strip_solvers_with_replicated_circles_all_reduce_their_corrections(world):
    res = _run(world, "circles")
    assert all(ok is True for _, ok, *_ in res), res


@pytest.mark.parametrize("world,case", [(2, "cut"), (3, "cut"), (2, "cut_ref")])
def test_bodies_cut_by_strip_edges_relax_their_cross_links_exactly(world, case):
    """N6: links whose ends are owned by neighbouring ranks (link.rs:18-27 relaxes any link between any two particles)"""
    res = _run(world, case)
    assert all(ok is True for _, ok, *_ in res), res


def test_nccl_strip_solvers_with_inverse_masses():
    res = _run(2, "kmass")
    assert all(ok is True for _, ok, *_ in res), res
    assert sum(r[2] for r in res) > 0
