#!/usr/bin/env python
"""One rank's share of the 8-GPU C5 run on a single GPU: the strip partition of rank `r` of `world`,
loaded as a strip solver WITHOUT neighbours (ghost slots stay empty), per-kernel CUDA-event times and the
graph-mode substep time.  Lets the per-rank kernels (grid window, scan variant, narrowphase) be tuned on
one GPU.  usage: strip_rank_kernels.py [label] [world=8] [rank=3]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bendy2d_b200 import scenes, strips

label = sys.argv[1] if len(sys.argv) > 1 else "default"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 3
# STRIP_RANK_SMALL=1: a 24-body field instead of the 16M scene (script self-test on the CPU emulator)
sc = scenes.c3_softbody_field(8, 3, 0, 0) if os.environ.get("STRIP_RANK_SMALL") == "1" else scenes.c5_softbody_field_16m()
sc.sub_steps, sc.dt = 8, float(np.float32(8 / 120.0))
part = strips.partition_scene(sc, world, None, sc.body_of)[rank]
sv = strips._load_part(part, -1)
sv.update(sc.dt, n=3)
sv.synchronize()
sv.timer_start()
sv.update(sc.dt, n=6)
graph_us = sv.timer_stop() * 1000 / (6 * 8)
sv.set_profiling(True)
sv.kernel_times(reset=True)
sv.update(sc.dt, n=2)
kt = sv.kernel_times(reset=True)
print(f"{label:44s} {sv.get_particle_len()} discs, grid {sv.grid()[3:5]}, stats {sv.stats()}: graph {graph_us:.1f} us/substep; eager: " +
      ", ".join(f"{k} {v['ms'] * 1000 / 16:.1f}us x{v['launches'] // 16}" for k, v in kt.items() if v["launches"]))
