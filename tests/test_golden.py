"""Golden-vector tests (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).

CPU: the oracle still reproduces the frozen vectors bit for bit.
GPU: the CUDA path reproduces the circle pile and the polygon heap bit for bit (both run in the
reference's own order on the device), and C1 in reference insertion order within 1e-5 after one step
(Gauss-Seidel order differs, see test_gpu_parity)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402
from helpers import bits  # noqa: E402

G = os.path.join(HERE, "golden")


def _same(a, b):
    return np.array_equal(bits(a), bits(b))


def test_oracle_reproduces_golden_vectors():
    g = np.load(os.path.join(G, "c1_reference_order.npz"))
    now = make_golden.c1_reference_order(int(g["n_updates"]))
    assert _same(now["pos"], g["pos"]) and _same(now["prev"], g["prev"]) and _same(now["circle_pos"], g["circle_pos"])
    g = np.load(os.path.join(G, "circle_pile.npz"))
    now = make_golden.circle_pile(int(g["n_updates"]))
    assert _same(now["pos"], g["pos"]) and _same(now["prev"], g["prev"])
    g = np.load(os.path.join(G, "polygon_heap.npz"))
    now = make_golden.polygon_heap(int(g["n_updates"]))
    for k in range(int(g["n_poly"])):
        assert _same(now[f"pos_{k}"], g[f"pos_{k}"]) and _same(now[f"center_{k}"], g[f"center_{k}"]), k


@pytest.mark.gpu
def test_gpu_reproduces_circle_pile_golden():
    from bendy2d_b200 import CircleLink, Link, Solver

    g = np.load(os.path.join(G, "circle_pile.npz"))
    s = Solver()
    s.bounds.size[:] = (30.0, 30.0)
    s.add_circles(g["init_pos"], g["radius"])
    s.add_circle_link(CircleLink(Link(0, 1, 4.0)))
    s.add_circle_link(CircleLink(Link(1, 5, 3.0)))
    s.update(1 / 120, n=int(g["n_updates"]))
    pos, prev, _ = s.read_circles()
    assert _same(pos, g["pos"]) and _same(prev, g["prev"])


@pytest.mark.gpu
def test_gpu_reproduces_polygon_heap_golden():
    from bendy2d_b200 import Polygon, Solver

    g = np.load(os.path.join(G, "polygon_heap.npz"))
    s = Solver()
    s.bounds.size[:] = (40.0, 40.0)
    for k in range(int(g["n_poly"])):
        s.add_polygon(Polygon.new(g[f"init_{k}"], bool(g["statics"][k])))
    s.update(1 / 120, n=int(g["n_updates"]))
    for k in range(int(g["n_poly"])):
        pos, prev, cen, _ = s.read_polygon(k)
        assert _same(pos, g[f"pos_{k}"]) and _same(prev, g[f"prev_{k}"]) and _same(cen, g[f"center_{k}"]), k
