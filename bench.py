#!/usr/bin/env python
"""bench.py — particle-substeps/s of the bendy2d solver substep on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--workload c1|c2|c3|c4|c5] [--impl reference]

A "step" is one frame = one Solver::update(dt) with sub_steps = 8 (8 substeps of 1/120 s; C1 uses
its own 1/60 s x 8).  N=1 runs C3, the 1M-particle softbody field the headline target is quoted on;
N>1 runs C5 (16M particles) sharded into spatial strips, one rank per GPU.
Timing: CUDA events on the solver's stream, per step, L2 flushed between steps (config.l2), max
over ranks.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SUBSTEPS_PER_STEP = 8


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
def make_scene(name: str):
    from bendy2d_b200 import scenes

    sc = {"c1": scenes.c1_softbody_blob, "c2": scenes.c2_free_particles, "c3": scenes.c3_softbody_field,
          "c4": scenes.c4_polygon_heavy, "c5": scenes.c5_softbody_field_16m}[name]()
    if name != "c1":  # one frame = 8 substeps of 1/120 (x8 and x0.125 are exact in f32)
        sc.dt = float(np.float32(np.float32(1.0 / 120.0) * np.float32(SUBSTEPS_PER_STEP)))
        sc.sub_steps = SUBSTEPS_PER_STEP
    return sc


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled under the benchmark's load (B200_PROFILING.md).  The timed region of
    the default run lasts about ten milliseconds - less than one nvidia-smi sample - so the sampler is started before
    the warm-up (nvidia-smi takes a moment to come up), the window that counts starts with the timed region and is
    kept open over untimed steps of the same workload that follow it until it is ~0.5 s long (window_s, extra_steps in
    the JSON); only samples whose own time stamp lies inside the window are used."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    WINDOW_S = 0.5

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    @staticmethod
    def _stamp(text, fallback):
        import datetime

        try:
            return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return fallback  # arrival time of the line

    def stop(self, t_begin: float, t_end: float, extra_steps: int = 0) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.08)  # let the last line of the window arrive
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for arrived, r in list(self.rows):
            try:
                ts = self._stamp(r[0], arrived)
                if not (t_begin <= ts <= t_end):
                    continue
                sm.append(float(r[2]))
                mx = float(r[3])
                for n, v in zip(names, r[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window_s": round(t_end - t_begin, 3), "extra_steps": extra_steps,
                "window": "the timed region + what follows it on this GPU: untimed steps of the same workload (one GPU) / "
                          "the sharded-vs-unsharded parity check (several)"}


def bind_to_gpu_numa_node(index: int):
    """One process per GPU: keep this rank's threads - and with them the first-touch placement of its pinned staging
    buffers - on the NUMA node its GPU hangs off, so that the e2e leg's host<->device copies do not cross the socket
    interconnect.  Best effort: any surprise (no sysfs, node -1, an empty intersection with the allowed CPUs) leaves
    the affinity alone.  Returns a short note for the log."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(index)  # the CUDA device itself (CUDA_VISIBLE_DEVICES may renumber)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0].strip().lower()
            bdf = out[-12:] if len(out) >= 12 else out  # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return f"GPU {index} ({bdf}): no NUMA node reported"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"GPU {index} ({bdf}): node {node} has none of the allowed CPUs"
        os.sched_setaffinity(0, cpus)
        return f"GPU {index} ({bdf}): bound to NUMA node {node} ({len(cpus)} CPUs)"
    except Exception as e:
        return f"GPU {index}: NUMA binding skipped ({type(e).__name__}: {e})"


# ------------------------------------------------------------------------------------------------
def cpu_baseline_sample(workload: str):
    """Bounded sample of the workload for the CPU oracle (BASELINE.md "CPU-baseline plan"): the scene at FULL SIZE
    wherever the oracle's substep is O(N) (C1, C2, C3; C5: one strip of eight, the share of one GPU of the 8-GPU
    run), stepped ONE substep of 1/120 s per step so that a --steps K --warmup W run stays within a minute.
    particle-substeps/s does not depend on how many substeps make a step."""
    from bendy2d_b200 import scenes, strips

    if workload == "c1":
        return scenes.c1_softbody_blob(), "C1 in full (400 particles, 1482 links, 1 circle), steps of 8 substeps"
    if workload == "c2":
        sc, what = scenes.c2_free_particles(), ("C2 at full size (100,000 discs); the oracle prunes pairs with a grid - the "
                                               "reference's literal all-pairs loop is O(n^2)")
    elif workload == "c4":
        sc, what = scenes.c4_polygon_heavy(20, 400), ("C4 at 8,000 discs / 400 polygons (of 200k / 10k): the reference's "
                                                     "polygon pass is all pairs, O(P^2)")
    elif workload == "c5":
        full = scenes.c5_softbody_field_16m()
        sc = strips.partition_scene(full, 8, None, full.body_of)[3].scene
        what = (f"C5: strip 3 of 8 at full size ({sc.n_particles:,} particles / {sc.n_links:,} links = one GPU's share of "
                "the 8-GPU run)")
    else:
        sc = scenes.c3_softbody_field()
        what = "C3 at full size (1,000,000 particles / 2,822,000 links / 200 circles / 500 polygons)"
    sc.dt, sc.sub_steps = float(np.float32(1.0 / 120.0)), 1
    return sc, what + ", steps of ONE substep of 1/120 s"


def oracle_for(sc):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_from_scene

    o = oracle_from_scene(sc)
    if sc.particle_radius > 0:
        h = 2.0 * sc.particle_radius
        o.set_grid(sc.bounds[0], sc.bounds[1], 1.0 / h, int(np.ceil(sc.bounds[2] / h)), int(np.ceil(sc.bounds[3] / h)))
    return o


def time_oracle(sc, steps: int, warmup: int):
    o = oracle_for(sc)
    for _ in range(warmup):
        o.update(sc.dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.update(sc.dt)
    dt = time.perf_counter() - t0
    return sc.n_points * sc.sub_steps * steps / dt, dt / steps * 1e3


def run_reference(args):
    """--impl reference: the reference's CPU path.  The Rust crate cannot be built in this image
    (no cargo/rustc, nalgebra not vendored) so this is the oracle port: single thread, because the
    reference is strictly single-threaded (plain for loops, no rayon: solver.rs:109-188)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload or ("c3" if args.gpus == 1 else "c5")
    sc, what = cpu_baseline_sample(workload)
    value, ms = time_oracle(sc, args.steps, args.warmup)
    full = make_scene(workload)
    line = {
        "impl": "reference", "metric": "particle-substeps/s", "value": value, "unit": "particle-substeps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": full.name, "substeps_per_step": sc.sub_steps, "sample": what,
                   "points_in_sample": sc.n_points, "points_in_workload": full.n_points},
        "cpu_baseline": {"value": value, "unit": "particle-substeps/s", "cores": 1, "kind": "port",
                         "sample": what, "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": "particle-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU: leave the CPU affinity of the ranks alone")
    ap.add_argument("--no-clock-window", action="store_true",
                    help="do not extend the nvidia-smi window with untimed steps (ncu launch lists)")
    ap.add_argument("--no-scaling-ref", action="store_true", help="skip the C5-on-one-GPU reference point")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps (not the headline)")
    ap.add_argument("--no-parity-check", action="store_true", help="N>1: skip the sharded-vs-unsharded bit compare")
    ap.add_argument("--grid-cell", type=float, default=0.0, help="override the broadphase cell edge (0 = auto)")
    ap.add_argument("--pack-points", type=int, default=0, help="override the link-partition pack target")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from bendy2d_b200 import Solver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bendy2d_b200 has no CPU path")
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa_bind else "not bound (single process)"
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    workload = args.workload or ("c3" if world == 1 else "c5")
    t_build = time.perf_counter()
    sc = make_scene(workload)
    n_points_total = sc.n_points
    n_warm = args.warmup  # honoured literally (the contract asks the caller for W >= 3)

    def fresh_solver():
        """every measured phase starts from the same state: initial scene + the warm-up steps"""
        if world > 1:
            from bendy2d_b200 import strips

            sv = strips.StripSolver(sc, rank, world, local, dist)
        else:
            sv = Solver(local)
            sc.load_into(sv)
            if args.grid_cell > 0:
                sv.set_grid_cell(args.grid_cell)
            if args.pack_points > 0:
                sv.set_plan_params(args.pack_points, 0)
        for _ in range(n_warm):
            sv.update(sc.dt)
        sv.synchronize()
        return sv

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # before the warm-up: it has to be running when the timed region begins
    solver = fresh_solver()
    info = solver.schedule_info()
    log(f"[rank {rank}] {numa_note}")
    log(f"[rank {rank}] scene {sc.name}: {sc.n_points} points, {sc.n_links} links, built in "
        f"{time.perf_counter() - t_build:.1f}s; schedule {info}")

    flush_buf = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        if flush_buf is not None:
            flush_buf.zero_()
            torch.cuda.synchronize()

    # ---- timed region: K steps, each bracketed by events on the solver's stream
    launches0 = solver.launch_count()
    barrier()
    clock_t0 = time.time()
    wall0 = time.perf_counter()
    total_ms = 0.0
    step_ms = []
    for _ in range(args.steps):
        flush_l2()
        if dist is not None:
            dist.barrier()
        solver.timer_start()
        solver.update(sc.dt)
        step_ms.append(solver.timer_stop())
        total_ms += step_ms[-1]
    barrier()
    wall = time.perf_counter() - wall0
    if world > 1:
        solver.check_halo()  # overflow / stale ownership would make the run invalid: raise
    gpu_launches = solver.launch_count() - launches0
    # ---- sharded runs: the final state of the timed run against the UNSHARDED 1-GPU run of the same updates
    # (C5 fits one GPU).  Puts the parity of the NCCL path into the bench record.
    parity_check = None
    if world > 1 and not args.no_parity_check:
        from bendy2d_b200 import strips

        lp, lq = solver.read_particles()
        gpos, gprev = strips.gather_global_state(dist, world, local, solver.part.global_index, lp, lq, sc.n_particles)
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from helpers import max_ulp

            one = Solver(local)
            sc.load_into(one)
            one.update(sc.dt, n=n_warm + args.steps)
            op, oq = one.read_particles()
            del one
            parity_check = {"max_ulp": max(max_ulp(gpos, op), max_ulp(gprev, oq)), "compared_particles": int(sc.n_particles),
                            "updates": n_warm + args.steps,
                            "what": "final pos and prev of the sharded timed run vs the unsharded 1-GPU run of the same updates"}
            log(f"[rank 0] parity_check {parity_check}")
        del gpos, gprev
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        lt = torch.tensor([gpu_launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        gpu_launches = int(lt.item())
    ms_per_step = total_ms / args.steps
    value = n_points_total * sc.sub_steps * args.steps / (total_ms * 1e-3)
    # ---- keep the clock window open under the same load (untimed; nothing below reads this solver's state again).
    # One GPU: further steps of the same workload, in pieces (the scene gets slower as it piles up), until the window
    # is long enough.  Several GPUs: no extra steps (every substep is a collective: ranks that disagreed on a count,
    # or one rank raising, would hang the others); there the window ends after the sharded-vs-unsharded parity check,
    # which steps the same scene on GPU 0.
    extra_steps = 0
    if world == 1 and not args.no_clock_window:
        budget = int(min(20000, ClockSampler.WINDOW_S * 1e3 / max(ms_per_step, 1e-3)))
        try:
            while extra_steps < budget and time.time() - clock_t0 < ClockSampler.WINDOW_S:
                n = min(32, budget - extra_steps)
                solver.update(sc.dt, n=n)
                solver.synchronize()
                extra_steps += n
        except Exception as e:  # the window is as long as it got
            log(f"[rank {rank}] clock window cut short: {e}")
    barrier()
    clocks = sampler.stop(clock_t0, time.time(), extra_steps) if rank == 0 else None

    # ---- warm-L2 figure (same steps back to back, one event pair) for context
    del solver
    solver = fresh_solver()
    barrier()
    solver.timer_start()
    solver.update(sc.dt, n=args.steps)
    warm_ms = solver.timer_stop()
    if dist is not None:
        t = torch.tensor([warm_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        warm_ms = float(t.item())
    value_warm = n_points_total * sc.sub_steps * args.steps / (warm_ms * 1e-3)

    # ---- e2e: host buffers in, host buffers out, every step (pinned memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        del solver
        solver = fresh_solver()
        n_local = solver.get_particle_len()
        h_pos = torch.empty((n_local, 2), dtype=torch.float32).pin_memory()
        h_prev = torch.empty((n_local, 2), dtype=torch.float32).pin_memory()
        np_pos, np_prev = h_pos.numpy(), h_prev.numpy()
        solver.read_particles(out_pos=np_pos, out_prev=np_prev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            solver.write_particles(np_pos, np_prev)          # H2D of the step's inputs
            solver.update(sc.dt)                             # 8 substeps
            solver.read_particles(out_pos=np_pos, out_prev=np_prev)  # D2H of the result (synchronises)
        barrier()
        e2e_s = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
            nb = torch.tensor([n_local], device="cuda", dtype=torch.int64)
            dist.all_reduce(nb)
            n_bytes = int(nb.item()) * 16
        else:
            n_bytes = n_local * 16
        e2e = {"value": n_points_total * sc.sub_steps * args.steps / e2e_s, "unit": "particle-substeps/s",
               "h2d_bytes_per_step": n_bytes, "d2h_bytes_per_step": n_bytes, "ms_per_step": e2e_s / args.steps * 1e3,
               "api": "bendy_write_particles + bendy_update + bendy_read_particles (pinned host buffers)"}

    # ---- per-kernel event timing (eager launches with event pairs) -> roofline of the dominant kernel
    del solver
    solver = fresh_solver()
    solver.set_profiling(True)
    solver.kernel_times(reset=True)
    prof_steps = max(2, min(args.steps, 25))
    for _ in range(prof_steps):
        flush_l2()
        solver.update(sc.dt)
    kt = solver.kernel_times(reset=True)
    solver.set_profiling(False)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    n_sub = prof_steps * sc.sub_steps
    grid = solver.grid() if sc.particle_radius > 0 else None
    shard = solver.local_scene() if world > 1 else sc
    alg = shard.algorithmic_bytes(grid)
    # kernel class -> algorithmic bytes per substep (SURVEY §8.4 per-unit figures x the units it processes).
    # With the disc grid on, the narrowphase launch ALSO does the particle-polygon contact (K4) and the
    # bounds+integrate (K1) of the free particles, so those bytes belong to it; the separate integrate
    # launch then only covers circle centres and polygon points.
    discs_on = shard.particle_radius > 0 and shard.n_particles > 0
    k1_particles = shard.n_particles * 32
    # the fused launch moves a particle's position ONCE: K4's "point R8 W8" (16 B/particle) is the same traffic as
    # K1's, so it is not counted a second time for this launch
    k4_tables = max(alg["K4_polygon"] - shard.n_particles * 16, 0)
    class_bytes = {"integrate": alg["K1_integrate"] - (k1_particles if discs_on else 0),
                   "links_local": alg["K3_links"], "links_global": 0, "grid_build": alg["K2_grid"],
                   "narrowphase": alg["K2_narrow"] + (k1_particles + k4_tables if discs_on else 0),
                   "poly_contact": 0 if discs_on else alg["K4_polygon"],
                   "fused": alg["K1_integrate"] + alg["K3_links"]}
    # BASELINE.md's contract figure: N_pts*32 + N_disc*(56+16) [N_cell = N_disc] + N_link*44, no K4 term: what the
    # ">= 50 % of the HBM roofline" target is quoted on (C3: 228 MB)
    n_disc = shard.n_particles if discs_on else 0
    contract_bytes = shard.n_points * 32 + n_disc * 72 + alg["K3_links"]
    kernels = {}
    total_k_ms = sum(v["ms"] for v in kt.values())
    for k, v in kt.items():
        if v["launches"]:
            per_sub = v["ms"] / n_sub
            kernels[k] = {"ms_per_substep": per_sub, "launches_per_substep": v["launches"] / n_sub,
                          "share": v["ms"] / total_k_ms if total_k_ms else 0.0}
            if class_bytes.get(k):
                kernels[k]["alg_GBps"] = class_bytes[k] / (per_sub * 1e-3) / 1e9
    peak, peak_src = peaks()
    dom = max((k for k in kernels if class_bytes.get(k)), key=lambda k: kernels[k]["ms_per_substep"])
    d = kernels[dom]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": d["alg_GBps"], "peak": peak, "unit": "GB/s",
                "frac": d["alg_GBps"] / peak, "traffic": None, "peak_source": peak_src,
                "avg_launch_ms": d["ms_per_substep"] / max(d["launches_per_substep"], 1e-9),
                "alg_bytes_per_substep_kernel": class_bytes[dom],
                "timing_note": ("avg_launch_ms and kernels{} are CUDA event pairs around eager launches of a separate run "
                                "(cold-ish caches, launch gaps included): they sum to more than the graph substep; shares only"),
                "substep": {"alg_bytes": contract_bytes,
                            "alg_bytes_note": "BASELINE.md contract: N_pts*32 + N_disc*72 + N_link*44 (per rank when sharded)",
                            "achieved": contract_bytes * sc.sub_steps / (ms_per_step * 1e-3) / 1e9,
                            "frac": contract_bytes * sc.sub_steps / (ms_per_step * 1e-3) / 1e9 / peak,
                            "frac_of_8TBps": contract_bytes * sc.sub_steps / (ms_per_step * 1e-3) / 1e9 / 8000.0,
                            "alg_bytes_full_formula": alg["total"],
                            "frac_full_formula": alg["total"] * sc.sub_steps / (ms_per_step * 1e-3) / 1e9 / peak,
                            "achieved_warm_l2": contract_bytes * sc.sub_steps * args.steps / (warm_ms * 1e-3) / 1e9},
                "kernels": kernels}
    # dram__bytes_read + dram__bytes_write of the dominant kernel from an `ncu --set full` capture: only valid for the
    # workload / shard it was captured on (profiles/traffic.json is keyed by workload, N = 1), else null
    ncu_traffic = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(ncu_traffic):
        try:
            roofline["traffic"] = json.load(open(ncu_traffic)).get(workload, {}).get(dom)
        except Exception:
            pass

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        csc, what = cpu_baseline_sample(workload)
        t0 = time.perf_counter()
        # ~10-20 s of single-core work (C3: 1.1 s per full-size substep)
        n_cpu_steps = {"c1": 250, "c2": 100, "c3": 12, "c4": 24, "c5": 5}[workload]
        v, _ = time_oracle(csc, n_cpu_steps, 1)
        cpu = {"value": v, "unit": "particle-substeps/s", "cores": 1, "kind": "port",
               "sample": f"{what}; {n_cpu_steps} steps after 1 warm-up",
               "host_cores_available": os.cpu_count(), "seconds": time.perf_counter() - t0}

    # ---- the 16M-particle scene on this one GPU: the N=1 point of the strong-scaling series that
    # `bench.py --gpus 2|4|8` continues (those runs shard C5; this run's headline stays C3)
    scaling_ref = None
    if world == 1 and workload == "c3" and not args.no_scaling_ref:
        del solver
        sc5 = make_scene("c5")
        s5 = Solver(local)
        sc5.load_into(s5)
        for _ in range(n_warm):  # the same window as the N >= 2 lines of the series
            s5.update(sc5.dt)
        s5.synchronize()
        n5, ms5 = args.steps, 0.0
        for _ in range(n5):
            flush_l2()
            s5.timer_start()
            s5.update(sc5.dt)
            ms5 += s5.timer_stop()
        scaling_ref = {"workload": sc5.name, "points": sc5.n_points, "n_gpus": 1, "steps": n5, "warmup": n_warm,
                       "ms_per_step": ms5 / n5, "value": sc5.n_points * sc5.sub_steps * n5 / (ms5 * 1e-3),
                       "unit": "particle-substeps/s"}
        del s5

    line = {
        "metric": "particle-substeps/s", "value": value, "unit": "particle-substeps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": sc.name, "points": n_points_total, "particles": sc.n_particles, "links": sc.n_links,
                   "circles": len(sc.circles_r), "polygons": len(sc.polygons), "substeps_per_step": sc.sub_steps,
                   "substep_dt": 1.0 / 120.0 if workload != "c1" else 1.0 / 480.0,
                   "l2": "warm between steps" if args.no_flush else "flushed between steps (256 MiB memset)",
                   "parallelism": f"strips{world}" if world > 1 else "single",
                   # the N=1 line's headline is C3 (the 1-GPU configuration); N>1 lines shard C5.  The N=1
                   # point of the C5 strong-scaling series is `strong_scaling_reference_c5_n1` of the N=1 line.
                   "series_note": ("C5 sharded over strips, total work fixed; compare with "
                                   "strong_scaling_reference_c5_n1 of the N=1 line, not with its C3 headline")
                   if world > 1 else
                   ("headline = C3 on one GPU; strong_scaling_reference_c5_n1 is the N=1 point of the C5 series "
                    "that --gpus 2/4/8 continue")},
        "value_warm_l2": value_warm, "wall_s_timed_region": wall,
        "ms_per_step_series": [round(x, 4) for x in step_ms],
        "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "parity_check": parity_check, "timed_substeps": args.steps * sc.sub_steps,
        "schedule": info, "strong_scaling_reference_c5_n1": scaling_ref,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
