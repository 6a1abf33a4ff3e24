"""Property tests of the host link planner (bendy_plan_links; no GPU): whatever the graph and the
packing parameters, the schedule must be a permutation of the links in which every (partition,
colour) bucket is vertex-disjoint — that is what makes the parallel relaxation arithmetically equal
to the sequential walk of the reference (solver.rs:144-146) over the exported order."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from bendy2d_b200.solver import BendyError, LinkPanic, plan_links

GLOBAL = 0xFFFFFFFF


def check_schedule(n, ab, pack, maxp):
    ab = np.asarray(ab, np.int64).reshape(-1, 2)
    m = len(ab)
    rank, perm, colour, part, info = plan_links(n, ab, pack, maxp)
    assert sorted(rank.tolist()) == list(range(n))
    assert sorted(perm.tolist()) == list(range(m))
    assert info["n_local_links"] + info["n_global_links"] == m
    assert info["n_local_colours"] <= 255
    eff_max = maxp if maxp else 4096
    r = rank.astype(np.int64)
    local = part != GLOBAL
    # 1. every bucket is an independent edge set
    key = np.where(local, part.astype(np.int64), -1) * (1 << 32) + colour
    for k in np.unique(key):
        ends = ab[key == k].ravel()
        assert len(np.unique(ends)) == len(ends)
    # 2. partitions are disjoint internal index ranges of at most max_points points
    spans = []
    for p in np.unique(part[local]):
        ends = r[ab[part == p].ravel()]
        spans.append((ends.min(), ends.max()))
        assert ends.max() - ends.min() < eff_max
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 < b0
    # 3. a global link really crosses two partitions' ranges (otherwise it should have been local), unless the
    # 255 local colours of its partition were used up at one of its ends (a hub)
    if spans and (~local).any():
        lo = np.array([s[0] for s in spans])
        deg = np.bincount(ab.ravel(), minlength=n)
        for a, b in ab[~local]:
            ia, ib = np.searchsorted(lo, r[a], "right"), np.searchsorted(lo, r[b], "right")
            assert ia != ib or max(deg[a], deg[b]) > 128 or not (spans[ia - 1][0] <= r[a] <= spans[ia - 1][1] and
                                                                 spans[ib - 1][0] <= r[b] <= spans[ib - 1][1])
    # 4. the exported order is partition-major, colour-major, global colours last
    is_glob = ~local[perm]
    assert (np.diff(is_glob.astype(int)) >= 0).all()
    assert (np.diff(key[perm][~is_glob]) >= 0).all()
    assert (np.diff(colour[perm][is_glob].astype(np.int64)) >= 0).all()
    # 5. unlinked points are numbered after all linked ones
    linked = np.zeros(n, bool)
    linked[ab.ravel()] = True
    if linked.any() and (~linked).any():
        assert r[~linked].min() > r[linked].max()
    return rank, perm, colour, part, info


@st.composite
def graphs(draw):
    n = draw(st.integers(2, 120))
    m = draw(st.integers(0, 300))
    shape = draw(st.sampled_from(["random", "chain", "stars", "cliques"]))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    if shape == "random":
        a, b = rng.integers(0, n, m), rng.integers(0, n, m)
    elif shape == "chain":
        a = rng.integers(0, n - 1, m)
        b = a + rng.integers(1, 3, m)
    elif shape == "stars":
        hubs = rng.integers(0, n, max(1, n // 20))
        a, b = rng.choice(hubs, m), rng.integers(0, n, m)
    else:
        size = draw(st.integers(2, 8))
        base = rng.integers(0, max(1, n // size), m) * size
        a, b = base + rng.integers(0, size, m), base + rng.integers(0, size, m)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    keep = (lo < hi) & (hi < n)
    ab = np.stack([lo[keep], hi[keep]], 1)
    pack = draw(st.sampled_from([0, 1, 8, 64]))
    maxp = draw(st.sampled_from([0, 16, 128]))
    if maxp and pack > maxp:
        pack = maxp
    return n, ab, pack, maxp


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(graphs())
def test_any_graph_yields_a_valid_schedule(g):
    n, ab, pack, maxp = g
    check_schedule(n, ab, pack, maxp)


@settings(max_examples=40, deadline=None)
@given(graphs())
def test_bucket_parallel_relaxation_equals_the_sequential_walk(g):
    """Relax the links with an arbitrary non-associative update, once sequentially in the exported order
    and once bucket by bucket from the pre-bucket state: identical floats, because buckets are disjoint."""
    n, ab, pack, maxp = g
    ab = np.asarray(ab, np.int64).reshape(-1, 2)
    if not len(ab):
        return
    rank, perm, colour, part, info = plan_links(n, ab, pack, maxp)
    rng = np.random.default_rng(1)
    x0 = rng.normal(size=n).astype(np.float32)

    def relax(xa, xb):
        d = np.float32(xa - xb)
        c = np.float32(d * np.float32(0.37)) + np.float32(0.01)
        return np.float32(xa - c), np.float32(xb + c)

    seq = x0.copy()
    for k in perm:
        a, b = ab[k]
        seq[a], seq[b] = relax(seq[a], seq[b])
    par = x0.copy()
    local = part != GLOBAL
    key = np.where(local, part.astype(np.int64), 1 << 40) * 1024 + colour
    order = perm[np.argsort(key[perm], kind="stable")]
    assert np.array_equal(order, perm)  # perm is already bucket-sorted
    for kk in np.unique(key):
        ks = perm[key[perm] == kk]
        before = par.copy()
        for k in ks:
            a, b = ab[k]
            par[a], par[b] = relax(before[a], before[b])
    assert np.array_equal(seq.view(np.uint32), par.view(np.uint32))


@given(graphs())
@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
def test_reference_order_levels_respect_insertion_order(g):
    """BENDY_LINKS_REFERENCE_ORDER: buckets are vertex-disjoint AND two links that share a point keep their
    insertion order in the exported sequence - so the schedule equals the reference's walk (solver.rs:144-146)."""
    n, ab, pack, maxp = g
    ab = np.asarray(ab, np.int64).reshape(-1, 2)
    rank, perm, colour, part, info = plan_links(n, ab, pack, maxp, reference_order=True)
    m = len(ab)
    assert sorted(perm.tolist()) == list(range(m))
    local = part != GLOBAL
    assert local.all() or (~local).all()  # levels per partition, or over the whole graph
    key = np.where(local, part.astype(np.int64), -1) * (1 << 32) + colour
    for k in np.unique(key):
        ends = ab[key == k].ravel()
        assert len(np.unique(ends)) == len(ends)
    where = np.empty(m, np.int64)
    where[perm] = np.arange(m)
    last = {}
    for k, (a, b) in enumerate(ab.tolist()):
        for v in (a, b):
            if v in last:
                assert where[last[v]] < where[k], "a later link that shares a point ran first"
            last[v] = k


def test_hubs_beyond_the_colour_tables_spill_into_extra_global_colours():
    # a star of degree d needs exactly d colours; the partition kernel's table holds 255 (kernels.cuh
    # K3_MAX_COLOURS), the cross-partition masks another 256; whatever is left gets one extra colour each.
    # The reference relaxes any graph (solver.rs:144-146), so the planner must never refuse one.
    for deg in (128, 255, 256, 400, 511, 512, 2000):
        ab = [[0, i] for i in range(1, deg + 1)]
        rank, perm, colour, part, info = check_schedule(deg + 1, ab, 0, 0)
        assert info["n_local_colours"] == min(deg, 255)
        assert info["n_global_links"] == max(0, deg - 255)
        assert info["n_global_colours"] == max(0, deg - 255)
    # two hubs sharing their leaves, in a partition that also holds ordinary links
    ab = [[0, i] for i in range(2, 700)] + [[1, i] for i in range(2, 700)] + [[i, i + 1] for i in range(2, 699)]
    check_schedule(700, ab, 0, 0)


def test_empty_and_degenerate_inputs():
    rank, perm, colour, part, info = plan_links(5, np.zeros((0, 2), np.uint32))
    assert sorted(rank.tolist()) == list(range(5)) and len(perm) == 0 and info["n_partitions"] == 0
    rank, perm, colour, part, info = plan_links(2, [[0, 1]])
    assert info["n_partitions"] == 1 and info["n_local_links"] == 1
    with pytest.raises(LinkPanic):
        plan_links(2, [[1, 0]])  # a < b (link.rs:6)
