#!/usr/bin/env python
"""Per-rank substep time and per-kernel event times of the strip path over NCCL, at the per-rank load of the 8-GPU C5
run (4000 bodies = 2M particles per rank), on however many GPUs torchrun was given:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/strips_timing.py [label]
The scene is C5's construction with 25*world x 160 bodies; the switches of libbendy2d_b200 come from the environment."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bendy2d_b200 import scenes, strips

label = sys.argv[1] if len(sys.argv) > 1 else "default"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = scenes.c5_softbody_field_16m(25 * world, 160)
sc.sub_steps, sc.dt = 8, float(np.float32(8 / 120.0))
sv = strips.StripSolver(sc, rank, world, local, dist)
sv.update(sc.dt, n=5)
sv.synchronize()
dist.barrier()
sv.timer_start()
sv.update(sc.dt, n=20)
ms = sv.timer_stop()
t = torch.tensor([ms], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
sv.check_halo()
sv.set_profiling(True)
sv.kernel_times(reset=True)
sv.update(sc.dt, n=2)
kt = sv.kernel_times(reset=True)
if rank in (0, world // 2):
    print(f"{label:40s} rank {rank}/{world}: {sv.get_particle_len()} discs; graph {float(t.item()) * 1000 / 160:.1f} us/substep "
          f"(max over ranks) = {sc.n_points * 160 / (float(t.item()) * 1e-3):.3e} particle-substeps/s; eager: " +
          ", ".join(f"{k} {v['ms'] * 1000 / 16:.1f}us x{v['launches'] // 16}" for k, v in kt.items() if v["launches"]), flush=True)
dist.destroy_process_group()
