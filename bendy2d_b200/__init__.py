"""bendy2d_b200 — B200-native (sm_100a) solver substep of the bendy2d 2D softbody engine.

Public surface = the reference's: Solver, Particle, Link, ParticleLink, CircleLink, Circle, Polygon,
Bounds (see solver.py).  Everything runs through libbendy2d_b200.so (C ABI: include/bendy2d_b200.h);
there is no CPU fallback.
"""
from .solver import (BendyError, Bounds, Circle, CircleLink, Link, LinkPanic, Particle, ParticleArray,
                     ParticleLink, Polygon, Solver, plan_links)

__all__ = ["BendyError", "Bounds", "Circle", "CircleLink", "Link", "LinkPanic", "Particle", "ParticleArray",
           "ParticleLink", "Polygon", "Solver", "plan_links"]
