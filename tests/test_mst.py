"""Minimum spanning tree / dendrogram / HDBSCAN.

CPU: the oracle (oracle/arborx_oracle.cpp: orc_mst, orc_dendrogram_union_find) against the reference's golden
vectors -- the equidistant / non-equidistant line cases of test/tstMinimumSpanningTree.cpp:90-140, the 1000-point
golden tree and the k = 5, 10, 15 total weights of test/tstMinimumSpanningTreeGoldenTest.cpp:86-150
(tests/golden/mst_golden.npz), the three dendrograms of test/tstDendrogram.cpp:56-113 -- and against an all-pairs
Kruskal.  GPU (-m gpu): the CUDA path through the C ABI against the oracle: identical edge sets and weights."""
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "mst_golden.npz")


def undirected(edges, weights):
    e = np.asarray(edges).reshape(-1, 2)
    lo, hi = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
    rows = sorted(zip(np.asarray(weights, np.float32).tolist(), lo.tolist(), hi.tolist()))
    return rows


def core_distances(pts, k):
    d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1, dtype=np.float32)).astype(np.float32)
    return np.sort(d, axis=1)[:, min(k, len(pts)) - 1]


def kruskal_total(pts, k):
    """total weight of an MST of the complete graph under the (mutual reachability) metric, float32 weights."""
    pts = np.asarray(pts, np.float32)
    n = len(pts)
    diff = pts[:, None, :] - pts[None, :, :]
    d = np.sqrt((diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]) + diff[..., 2] * diff[..., 2]).astype(np.float32)
    if k > 1:
        c = np.sort(d, axis=1)[:, min(k, n) - 1]
        d = np.maximum(np.maximum(c[:, None], c[None, :]), d)
    iu = np.triu_indices(n, 1)
    w = d[iu]
    order = np.argsort(w, kind="stable")
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    total, used, ws = 0.0, 0, []
    for e in order:
        a, b = find(int(iu[0][e])), find(int(iu[1][e]))
        if a != b:
            parent[a] = b
            ws.append(w[e])
            used += 1
            if used == n - 1:
                break
    return np.sort(np.array(ws, np.float32))


def reorder_to_weight_order(weights, parents):
    """tstDendrogram.cpp:141-200: renumber a dendrogram whose edges are in the hybrid algorithm's order into the
    ascending-weight order the UNION_FIND dendrogram uses (weights must be distinct for the two to be comparable)."""
    m = len(weights)
    order = np.argsort(weights, kind="stable")
    rev = np.empty(m, np.int64)
    rev[order] = np.arange(m)
    out = np.empty_like(parents)
    pe = parents[:m]
    out[:m][rev] = np.where(pe >= 0, rev[np.maximum(pe, 0)], -1)
    out[m:] = rev[parents[m:]]
    return out, weights[order]


def check_dendrogram(edges, weights, parents, heights):
    """A dendrogram of a spanning tree, ties or not: one root, every edge node has exactly two children, heights never
    decrease towards the root, heights are the edge weights, and the two end points of every tree edge meet at the
    height of that edge (single linkage).  With distinct weights they meet at the edge itself; among edges of EQUAL
    weight in one chain the hybrid algorithm leaves the order open (BoruvkaHelpers.hpp:603-622), so only the height
    is required there."""
    m = len(weights)
    n = m + 1
    assert parents.shape == (2 * m + 1,) and np.array_equal(heights, weights)
    assert int((parents[:m] == -1).sum()) == 1 and parents[m:].min() >= 0 and parents.max() < m
    children = np.bincount(parents[parents >= 0], minlength=m)
    assert np.array_equal(children, np.full(m, 2))
    pe = parents[:m]
    has = pe >= 0
    assert (weights[pe[has]] >= weights[has]).all()
    # depth of every node, then the LCA of each edge's end points by walking up
    depth = np.full(2 * m + 1, -1, np.int64)
    root = int(np.nonzero(pe == -1)[0][0])
    depth[root] = 0
    order = np.argsort(-weights, kind="stable")
    for _ in range(4 * m + 8):
        todo = np.nonzero(depth < 0)[0]
        if len(todo) == 0:
            break
        ready = todo[depth[parents[todo]] >= 0]
        depth[ready] = depth[parents[ready]] + 1
    assert (depth >= 0).all()
    for e in range(0, m, max(1, m // 2000)):
        a, b = int(edges[e, 0]) + m, int(edges[e, 1]) + m
        while a != b:
            if depth[a] >= depth[b]:
                a = int(parents[a])
            else:
                b = int(parents[b])
        assert weights[a] == weights[e], (e, a)


LINE = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [4, 0, 0]], np.float32)
# tstMinimumSpanningTree.cpp:98-124
LINE_REF = {
    1: [(0, 1, 1), (1, 2, 1), (2, 3, 1), (3, 4, 1)],
    2: [(0, 1, 1), (1, 2, 1), (2, 3, 1), (3, 4, 1)],
    3: [(0, 1, 2), (1, 2, 1), (2, 3, 1), (2, 4, 2)],
    4: [(0, 1, 3), (1, 2, 2), (1, 3, 2), (1, 4, 3)],
    5: [(0, 1, 4), (1, 2, 3), (1, 3, 3), (0, 4, 4)],
}
UNEVEN = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [6, 0, 0], [10, 0, 0]], np.float32)
# :126-139
UNEVEN_REF = {k: [(0, 1, 1), (1, 2, 1), (2, 3, 1), (3, 4, 3), (4, 5, 4)] for k in (1, 2)}

# tstDendrogram.cpp:56-113: (edges, parents, heights)
DENDROGRAMS = [
    ([(0, 1, 3.0)], [-1, 0, 0], [3.0]),
    ([(0, 3, 7.0), (1, 2, 3.0), (0, 1, 2.0)], [1, 2, -1, 0, 0, 1, 2], [2.0, 3.0, 7.0]),
    ([(2, 3, 2.0), (2, 0, 9.0), (0, 1, 3.0)], [2, 2, -1, 1, 1, 0, 0], [2.0, 3.0, 9.0]),
]


def ref_rows(ref):
    return sorted((float(w), min(a, b), max(a, b)) for a, b, w in ref)


def check_spanning(n, edges):
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b in np.asarray(edges).tolist():
        ra, rb = find(a), find(b)
        assert ra != rb, "cycle"
        parent[ra] = rb
    assert len({find(i) for i in range(n)}) == 1


# ------------------------------------------------------------------ oracle (CPU) ----
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_oracle_line_cases(k):
    e, w = oracle.mst(LINE, k)
    assert undirected(e, w) == ref_rows(LINE_REF[k])
    if k in UNEVEN_REF:
        e, w = oracle.mst(UNEVEN, k)
        assert undirected(e, w) == ref_rows(UNEVEN_REF[k])


def test_oracle_golden_tree():
    g = np.load(GOLDEN)
    pts = g["points"].astype(np.float32)
    e, w = oracle.mst(pts, 1)
    got = {(min(a, b), max(a, b)) for a, b in e.tolist()}
    ref = {(min(a, b), max(a, b)) for a, b in g["edges"].tolist()}
    assert got == ref
    # the reference's test compares the vertex pairs only (UndirectedEdge::operator==, :27-33); the csv weights are
    # double-precision distances of the csv coordinates, the tree's are float distances of the float coordinates
    gw = dict(zip(map(lambda r: (min(r), max(r)), g["edges"].tolist()), g["weights"].tolist()))
    for (a, b), x in zip(e.tolist(), w.tolist()):
        assert abs(gw[(min(a, b), max(a, b))] - x) <= 2e-5 * x
    # tstMinimumSpanningTreeGoldenTest.cpp:126-150: total weights for k = 5, 10, 15, relative tolerance 1e-8
    for k, total in zip(g["total_weight_k"].tolist(), g["total_weight"].tolist()):
        e, w = oracle.mst(pts, k)
        assert abs(float(np.sum(w.astype(np.float64))) - total) <= 1e-8 * total


@pytest.mark.parametrize("k", [1, 3, 8])
def test_oracle_vs_kruskal(k):
    rng = np.random.default_rng(7 + k)
    for pts in (rng.random((600, 3), dtype=np.float32),
                # a lattice: many equal weights, the tree is unique only through the pair order
                np.stack(np.meshgrid(*[np.arange(7, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)):
        e, w = oracle.mst(pts, k)
        check_spanning(len(pts), e)
        assert np.array_equal(np.sort(w), kruskal_total(pts, k))
        # every edge carries its own metric value
        c = core_distances(pts, k) if k > 1 else np.zeros(len(pts), np.float32)
        diff = pts[e[:, 0]] - pts[e[:, 1]]
        d = np.sqrt((diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]).astype(np.float32)
        assert np.array_equal(np.maximum(np.maximum(c[e[:, 0]], c[e[:, 1]]), d), w)


def test_oracle_dendrogram_golden():
    for edges, parents, heights in DENDROGRAMS:
        e = np.array([(a, b) for a, b, _ in edges], np.int32)
        w = np.array([x for _, _, x in edges], np.float32)
        p, h = oracle.dendrogram(e, w)
        assert p.tolist() == parents and h.tolist() == heights


def test_oracle_degenerate():
    e, w = oracle.mst(np.zeros((1, 3), np.float32), 1)
    assert e.shape == (0, 2) and w.shape == (0,)
    e, w = oracle.mst(np.array([[0, 0, 0], [3, 4, 0]], np.float32), 1)
    assert undirected(e, w) == [(5.0, 0, 1)]
    # duplicates: zero-weight edges, still a spanning tree
    pts = np.repeat(np.random.default_rng(3).random((40, 3), dtype=np.float32), 3, axis=0)
    e, w = oracle.mst(pts, 1)
    check_spanning(len(pts), e)
    assert (w == 0).sum() == 80


@pytest.mark.parametrize("n,k", [(2, 1), (7, 2), (50, 1), (3000, 1), (3000, 5), (20000, 3)])
def test_oracle_hybrid_dendrogram(n, k):
    """BoruvkaMode::HDBSCAN in the oracle: same tree as MST mode, a valid dendrogram, and for distinct weights exactly
    the UNION_FIND dendrogram after renumbering (the reference's own check, tstDendrogram.cpp:115-200)."""
    rng = np.random.default_rng(100 + n + k)
    pts = (rng.random((n, 3), dtype=np.float32) * 100).astype(np.float32)
    e, w, p, h = oracle.mst_hdbscan(pts, k)
    e0, w0 = oracle.mst(pts, k)
    assert undirected(e, w) == undirected(e0, w0)
    check_dendrogram(e, w, p, h)
    if len(np.unique(w)) == len(w):
        p2, h2 = reorder_to_weight_order(w, p)
        pu, hu = oracle.dendrogram(e, w)
        assert np.array_equal(p2, pu) and np.array_equal(h2, hu)


def test_oracle_hybrid_dendrogram_ties():
    lattice = np.stack(np.meshgrid(*[np.arange(9, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
    for k in (1, 4):
        e, w, p, h = oracle.mst_hdbscan(lattice, k)
        check_dendrogram(e, w, p, h)
    e, w, p, h = oracle.mst_hdbscan(LINE, 3)
    check_dendrogram(e, w, p, h)
    e, w, p, h = oracle.mst_hdbscan(np.zeros((1, 3), np.float32), 2)
    assert p.tolist() == [-1]


# tstDetailsMutualReachabilityDistance.cpp:97-128: core distances (distance to the k-th nearest point, the point itself
# included) of the two line clouds
CORE_LINE = {1: [0, 0, 0, 0, 0], 2: [1, 1, 1, 1, 1], 3: [2, 1, 1, 1, 2], 4: [3, 2, 2, 2, 3], 5: [4, 3, 2, 3, 4]}
CORE_UNEVEN = {2: [1, 1, 1, 1, 3, 4], 3: [2, 1, 1, 2, 4, 7], 4: [3, 2, 2, 3, 4, 8], 5: [6, 5, 4, 3, 5, 9],
               6: [10, 9, 8, 7, 6, 10]}


def test_oracle_core_distances_golden():
    for pts, table in ((LINE, CORE_LINE), (UNEVEN, CORE_UNEVEN)):
        tree = oracle.Tree(pts)
        for k, ref in table.items():
            off, idx, d = tree.nearest_crs(pts, k)
            assert np.array_equal(np.diff(off), np.full(len(pts), k))
            assert d.reshape(len(pts), k)[:, -1].tolist() == [float(x) for x in ref]
            assert core_distances(pts, k).tolist() == [float(x) for x in ref]


def test_oracle_fuzz_small_clouds():
    """Random small clouds -- uniform, a 4^3 integer lattice with repeated points, duplicated points -- and k = 1...7:
    the oracle's tree weighs what an all-pairs Kruskal weighs, and its hybrid dendrogram is a valid one."""
    rng = np.random.default_rng(2024)
    for trial in range(45):
        n, k = int(rng.integers(2, 300)), int(rng.integers(1, 8))
        if trial % 3 == 0:
            pts = rng.random((n, 3), dtype=np.float32)
        elif trial % 3 == 1:
            pts = rng.integers(0, 4, (n, 3)).astype(np.float32)
        else:
            pts = np.repeat(rng.random((max(n // 2, 1), 3), dtype=np.float32), 2, axis=0)
        e, w = oracle.mst(pts, k)
        check_spanning(len(pts), e)
        assert np.array_equal(np.sort(w), kruskal_total(pts, k)), (trial, n, k)
        e2, w2, p2, h2 = oracle.mst_hdbscan(pts, k)
        assert np.array_equal(np.sort(w2), np.sort(w))
        check_dendrogram(e2, w2, p2, h2)


# ------------------------------------------------------------------ CUDA path ----
def _gpu():
    import torch

    import arborx_b200 as abx
    return torch, abx, abx.ExecutionSpace()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2, 3, 4, 5])
def test_cuda_line_cases(k):
    torch, abx, space = _gpu()
    mst = abx.MinimumSpanningTree(space, torch.from_numpy(LINE).cuda(), k)
    space.fence()
    assert undirected(mst.edges.cpu().numpy(), mst.weights.cpu().numpy()) == ref_rows(LINE_REF[k])
    if k in UNEVEN_REF:
        mst = abx.MinimumSpanningTree(space, torch.from_numpy(UNEVEN), k)  # host entry point
        assert undirected(mst.edges.numpy(), mst.weights.numpy()) == ref_rows(UNEVEN_REF[k])


@pytest.mark.gpu
def test_cuda_golden_tree():
    torch, abx, space = _gpu()
    g = np.load(GOLDEN)
    pts = g["points"].astype(np.float32)
    mst = abx.MinimumSpanningTree(space, torch.from_numpy(pts).cuda())
    space.fence()
    got = {(min(a, b), max(a, b)) for a, b in mst.edges.cpu().numpy().tolist()}
    assert got == {(min(a, b), max(a, b)) for a, b in g["edges"].tolist()}
    for k, total in zip(g["total_weight_k"].tolist(), g["total_weight"].tolist()):
        mst = abx.MinimumSpanningTree(space, torch.from_numpy(pts).cuda(), k)
        space.fence()
        assert abs(float(mst.weights.double().sum().item()) - total) <= 1e-8 * total


@pytest.mark.gpu
@pytest.mark.parametrize("n,k,kind", [(2, 1, "random"), (3, 2, "random"), (1000, 1, "random"), (5000, 4, "random"),
                                      (343, 1, "lattice"), (343, 5, "lattice"), (4096, 2, "lattice"),
                                      (3000, 1, "duplicates"), (3000, 6, "duplicates"), (200_000, 1, "random"),
                                      (200_000, 5, "clustered"), (50, 80, "random")])
def test_cuda_vs_oracle(n, k, kind):
    torch, abx, space = _gpu()
    rng = np.random.default_rng(n * 31 + k)
    if kind == "lattice":
        m = round(n ** (1 / 3))
        pts = np.stack(np.meshgrid(*[np.arange(m, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
    elif kind == "duplicates":
        pts = np.repeat(rng.random((n // 3, 3), dtype=np.float32), 3, axis=0)
    elif kind == "clustered":
        from tests import clouds
        pts = clouds.gan_tao(3, n)
    else:
        pts = rng.random((n, 3), dtype=np.float32) * 100
    pts = np.ascontiguousarray(pts, np.float32)
    mst = abx.MinimumSpanningTree(space, torch.from_numpy(pts).cuda(), k)
    space.fence()
    e, w = oracle.mst(pts, k)
    assert undirected(mst.edges.cpu().numpy(), mst.weights.cpu().numpy()) == undirected(e, w)
    assert mst.iterations >= 1


@pytest.mark.gpu
def test_cuda_dendrogram_and_hdbscan():
    torch, abx, space = _gpu()
    for edges, parents, heights in DENDROGRAMS:
        e = torch.tensor([(a, b) for a, b, _ in edges], dtype=torch.int32).cuda()
        w = torch.tensor([x for _, _, x in edges], dtype=torch.float32).cuda()
        d = abx.Dendrogram(space, e, w)
        space.fence()
        assert d._parents.cpu().tolist() == parents and d._parent_heights.cpu().tolist() == heights
    # hdbscan = MST(core_min_size) + dendrogram; distinct weights make the dendrogram unique
    rng = np.random.default_rng(11)
    pts = (rng.random((3000, 3), dtype=np.float32) * 100).astype(np.float32)
    for k in (1, 5):
        d = abx.hdbscan(space, torch.from_numpy(pts).cuda(), k, abx.DENDROGRAM_UNION_FIND)
        space.fence()
        e, w = oracle.mst(pts, k)
        if len(np.unique(w)) != len(w):
            continue
        p, h = oracle.dendrogram(e, w)
        assert np.array_equal(d._parents.cpu().numpy(), p)
        assert np.array_equal(d._parent_heights.cpu().numpy(), h)
    # single vertex
    for impl in (abx.DENDROGRAM_BORUVKA, abx.DENDROGRAM_UNION_FIND):
        d = abx.hdbscan(space, torch.zeros((1, 3), device="cuda"), 2, impl)
        space.fence()
        assert d._parents.cpu().tolist() == [-1]


@pytest.mark.gpu
@pytest.mark.parametrize("n,k,kind", [(2, 1, "random"), (7, 2, "random"), (3000, 1, "random"), (3000, 5, "random"),
                                      (729, 1, "lattice"), (729, 4, "lattice"), (3000, 1, "duplicates"),
                                      (200_000, 1, "random"), (200_000, 5, "clustered")])
def test_cuda_hybrid_dendrogram(n, k, kind):
    """MinimumSpanningTree in HDBSCAN mode on the device (no host pass): the same tree, a valid dendrogram, and for
    distinct weights exactly the UNION_FIND dendrogram (and the oracle's hybrid dendrogram) after renumbering."""
    torch, abx, space = _gpu()
    rng = np.random.default_rng(n * 17 + k)
    if kind == "lattice":
        m = round(n ** (1 / 3))
        pts = np.stack(np.meshgrid(*[np.arange(m, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
    elif kind == "duplicates":
        pts = np.repeat(rng.random((n // 3, 3), dtype=np.float32), 3, axis=0)
    elif kind == "clustered":
        from tests import clouds
        pts = clouds.gan_tao(3, n)
    else:
        pts = rng.random((n, 3), dtype=np.float32) * 100
    pts = np.ascontiguousarray(pts, np.float32)
    mst = abx.MinimumSpanningTree(space, torch.from_numpy(pts).cuda(), k, mode="hdbscan")
    space.fence()
    e, w = mst.edges.cpu().numpy(), mst.weights.cpu().numpy()
    p, h = mst.dendrogram_parents.cpu().numpy(), mst.dendrogram_parent_heights.cpu().numpy()
    e0, w0 = oracle.mst(pts, k)
    assert undirected(e, w) == undirected(e0, w0)
    check_dendrogram(e, w, p, h)
    if len(np.unique(w)) == len(w):
        p2, h2 = reorder_to_weight_order(w, p)
        pu, hu = oracle.dendrogram(e, w)
        assert np.array_equal(p2, pu) and np.array_equal(h2, hu)
        # (the chains themselves are numbered by the order in which components appended their edges, which is not
        # deterministic on the device: only the renumbered form is comparable, also with the oracle's hybrid result)
        eo, wo, po, ho = oracle.mst_hdbscan(pts, k)
        p3, h3 = reorder_to_weight_order(wo, po)
        assert np.array_equal(p2, p3) and np.array_equal(h2, h3)
    d = abx.hdbscan(space, torch.from_numpy(pts).cuda(), k)  # BORUVKA is the default, as in the reference
    space.fence()
    dp, dh = d._parents.cpu().numpy(), d._parent_heights.cpu().numpy()
    assert np.array_equal(np.sort(dh), np.sort(w))
    if len(np.unique(w)) == len(w):  # two runs number their chains differently: compare the renumbered forms
        p4, h4 = reorder_to_weight_order(dh, dp)
        assert np.array_equal(p4, p2) and np.array_equal(h4, h2)


@pytest.mark.gpu
def test_cuda_errors():
    torch, abx, space = _gpu()
    with pytest.raises(Exception):
        abx.MinimumSpanningTree(space, torch.zeros((4, 3), device="cuda"), 0)
    mst = abx.MinimumSpanningTree(space, torch.zeros((1, 3), device="cuda"))
    assert mst.edges.shape == (0, 2)
    mst = abx.MinimumSpanningTree(space, torch.zeros((0, 3), device="cuda"))
    assert mst.edges.shape == (0, 2)
