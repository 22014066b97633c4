"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on identical seeded inputs.  Integer / index results must be bit-exact
(sorted codes, permutation, node arrays, per-row index sets, core labels); kNN
distances within 1e-6 relative (BASELINE.json north_star) -- in fact they are
compared bit-exactly against the oracle and 1e-6 against brute force."""
import numpy as np
import pytest

import oracle
from tests import brute, clouds, dbscan_checks
from tests.engines import (PRED_BOX, PRED_POINT, PRED_SPHERE, PRIM_BOX, PRIM_POINT, PRIM_TRI, CudaEngineLazy,
                           OracleEngine, rows_of)

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def cuda():
    e = CudaEngineLazy()
    e.ensure()
    return e


@pytest.fixture(scope="module")
def orc():
    return OracleEngine()


def boxes_from(seed, n, scale=0.02):
    lo = clouds.uniform01(seed, n)
    ext = clouds.uniform01(seed + 11, n) * F(scale)
    return np.concatenate([lo, lo + ext], 1).astype(F)


def tris_from(seed, n, scale=0.03):
    a = clouds.uniform01(seed, n)
    b = a + (clouds.uniform01(seed + 21, n) - F(0.5)) * F(scale)
    c = a + (clouds.uniform01(seed + 22, n) - F(0.5)) * F(scale)
    return np.concatenate([a, b, c], 1).astype(F)


# ---------------------------------------------------------------- sort / scan ----
@pytest.mark.parametrize("n", [1, 2, 31, 255, 4095, 4096, 4097, 100_003, 1_500_001])
def test_sort_u64(cuda, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    if n > 100:
        keys[::7] = keys[3]  # heavy duplicates exercise stability
    k, p = cuda.sort_u64(keys)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(p, order.astype(np.uint32))


@pytest.mark.parametrize("n", [2, 300, 2047, 2048, 2049, 4097, 100_003, 3_000_001])
@pytest.mark.parametrize("run", [0, 3, 200, 256, 257, 700])
def test_sort_u64_prefix_runs(cuda, n, run):
    """Top-digit passes + segment fix-up (abx_sort.cu): unique-ish keys take the fix-up path; runs of
    keys sharing their top bits up to and beyond its 256-key limit exercise the escalation steps."""
    rng = np.random.default_rng(n * 1000 + run)
    keys = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
    if run and n > run:
        # runs sharing the top 24 bits (low bits random, some exact duplicates), scattered over the input,
        # one of them planted so that it straddles a fix-up tile boundary after sorting
        for r in range(max(1, min(40, n // (4 * run)))):
            pos = rng.choice(n, run, replace=False)
            pre = np.uint64(0) if r == 0 else (keys[pos[0]] >> np.uint64(39)) << np.uint64(39)
            low = rng.integers(0, 2 ** 39, run, dtype=np.uint64)
            low[::5] = low[0]
            keys[pos] = pre | low
    k, p = cuda.sort_u64(keys)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(p, order.astype(np.uint32))
    assert np.array_equal(k, keys[order])


@pytest.mark.parametrize("n", [2048, 2049, 16_384, 16_385, 524_288, 524_289, 4_194_305])
def test_sort_u64_level_boundaries(cuda, n):
    """Sizes on both sides of the points where the number of top digits of the fix-up path changes
    (n / 8 against 2^8, 2^16, 2^24; abx_sort.cu sortPairsDB), keys over all 64 bits."""
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2 ** 64, n, dtype=np.uint64)
    keys[n // 3] = keys[n // 5]
    k, p = cuda.sort_u64(keys)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(p, order.astype(np.uint32))
    assert np.array_equal(k, keys[order])


@pytest.mark.parametrize("distinct", [50, 1000, 40_000])
def test_sort_u64_clustered_prefixes(cuda, distinct):
    """Most keys share few 40-bit prefixes (the Morton codes of a clustered cloud): the prefix sample sends the
    sort to the plain LSD (or the 5-digit level) instead of a fix-up that would overflow."""
    n = 600_000
    rng = np.random.default_rng(distinct)
    pre = rng.integers(0, 2 ** 40, distinct, dtype=np.uint64) << np.uint64(23)
    keys = pre[rng.integers(0, distinct, n)] | rng.integers(0, 2 ** 23, n, dtype=np.uint64)
    keys[:1000] = rng.integers(0, 2 ** 63, 1000, dtype=np.uint64)  # a little noise that spans the range
    k, p = cuda.sort_u64(keys)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(p, order.astype(np.uint32))
    assert np.array_equal(k, keys[order])


@pytest.mark.parametrize("n", [1, 5, 4096, 300_001])
def test_sort_u32(cuda, n):
    rng = np.random.default_rng(n + 1)
    keys = rng.integers(0, 2 ** 30, n, dtype=np.uint32)
    keys[::3] = 12345
    k, p = cuda.sort_u32(keys)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(p, order.astype(np.uint32))


def test_sort_low_entropy(cuda):
    keys = np.zeros(50_000, np.uint64)
    keys[25_000:] = 1 << 62
    k, p = cuda.sort_u64(keys[::-1].copy())
    assert np.all(np.diff(k.astype(np.float64)) >= 0)
    assert np.array_equal(p[:25_000], np.arange(25_000, 50_000, dtype=np.uint32))


@pytest.mark.parametrize("n", [1, 2, 2048, 2049, 1_000_003])
def test_exclusive_scan(cuda, n):
    import ctypes as C
    import torch
    from arborx_b200 import _lib
    rng = np.random.default_rng(n)
    v = rng.integers(0, 50, n + 1).astype(np.int32)
    t = torch.as_tensor(v).cuda()
    out = torch.empty_like(t)
    _lib.check(_lib.lib().abx_exclusive_scan_i32(cuda.space.handle, C.c_void_p(t.data_ptr()),
                                                 C.c_void_p(out.data_ptr()), n + 1))
    ref = np.concatenate([[0], np.cumsum(v[:-1])]).astype(np.int32)
    assert np.array_equal(out.cpu().numpy(), ref)
    # in place
    _lib.check(_lib.lib().abx_exclusive_scan_i32(cuda.space.handle, C.c_void_p(t.data_ptr()),
                                                 C.c_void_p(t.data_ptr()), n + 1))
    assert np.array_equal(t.cpu().numpy(), ref)


# ------------------------------------------------------- bounds / codes / tree ----
PRIM_CASES = [("points", PRIM_POINT, lambda s, n: clouds.filled_box(s, n)),
              ("boxes", PRIM_BOX, boxes_from), ("triangles", PRIM_TRI, tris_from)]


@pytest.mark.parametrize("name,kind,gen", PRIM_CASES, ids=[c[0] for c in PRIM_CASES])
@pytest.mark.parametrize("n", [2, 3, 1000, 200_000])
def test_bounds_codes_tree_structure(cuda, orc, name, kind, gen, n):
    prims = gen(42, n)
    b = cuda.scene_bounds(prims, kind)
    assert np.array_equal(b, orc.scene_bounds(prims, kind))
    assert np.array_equal(cuda.morton64_codes(prims, b, kind), orc.morton64_codes(prims, b, kind))
    tc, to = cuda.build(prims, kind), orc.build(prims, kind)
    assert np.array_equal(tc.bounds(), to.bounds())
    dc, do = tc.export(), to.export()
    for key in ("codes", "leaf_index", "leaf_rope", "left_child", "rope", "boxes"):
        assert np.array_equal(dc[key], do[key]), key


def test_tree_structure_duplicates(cuda, orc):
    # many identical points and identical codes: index-bit fallback of delta()
    rng = np.random.default_rng(3)
    base = rng.random((50, 3)).astype(F)
    pts = base[rng.integers(0, 50, 20_000)]
    tc, to = cuda.build(pts, PRIM_POINT), orc.build(pts, PRIM_POINT)
    dc, do = tc.export(), to.export()
    for key in ("codes", "leaf_index", "leaf_rope", "left_child", "rope", "boxes"):
        assert np.array_equal(dc[key], do[key]), key


def test_tree_structure_1m(cuda, orc):
    pts = clouds.filled_box(0x5EED0001, 1_000_000)
    tc, to = cuda.build(pts, PRIM_POINT), orc.build(pts, PRIM_POINT)
    dc, do = tc.export(), to.export()
    for key in ("codes", "leaf_index", "leaf_rope", "left_child", "rope", "boxes"):
        assert np.array_equal(dc[key], do[key]), key


# -------------------------------------------------------------- spatial queries ----
def same_rows(a, b):
    (oa, ia), (ob, ib) = a, b
    assert np.array_equal(oa, ob)
    # sort inside rows: segment-wise sort via lexsort on (value, row)
    def canon(off, idx):
        row = np.repeat(np.arange(len(off) - 1), np.diff(off))
        order = np.lexsort((idx, row))
        return idx[order]
    assert np.array_equal(canon(oa, ia), canon(ob, ib))


@pytest.mark.parametrize("name,kind,gen", PRIM_CASES, ids=[c[0] for c in PRIM_CASES])
@pytest.mark.parametrize("sort_predicates", [True, False])
def test_spatial_sphere_vs_oracle(cuda, orc, name, kind, gen, sort_predicates):
    n, q = 50_000, 20_000
    prims = gen(7, n)
    if kind == PRIM_POINT:
        prims = (prims / F(np.cbrt(n)) * F(0.5) + F(0.5)).astype(F)
    centers = clouds.uniform01(99, q)
    radii = (clouds.uniform01(98, q, 1) * F(0.06)).astype(F)
    spheres = np.concatenate([centers, radii], 1).astype(F)
    tc, to = cuda.build(prims, kind), orc.build(prims, kind)
    same_rows(tc.spatial_crs(spheres, PRED_SPHERE, sort_predicates), to.spatial_crs(spheres, PRED_SPHERE))
    cnt = tc.spatial_count(spheres, PRED_SPHERE)
    assert np.array_equal(cnt, to.spatial_count(spheres, PRED_SPHERE))
    lim = tc.spatial_count(spheres, PRED_SPHERE, limit=3)
    assert np.array_equal(lim, np.minimum(cnt, 3))


@pytest.mark.parametrize("kind,gen", [(PRIM_POINT, lambda s, n: clouds.uniform01(s, n)), (PRIM_BOX, boxes_from)],
                         ids=["points", "boxes"])
def test_spatial_box_and_point_predicates_vs_oracle(cuda, orc, kind, gen):
    n, q = 30_000, 10_000
    prims = gen(5, n)
    qb = boxes_from(77, q, 0.08)
    tc, to = cuda.build(prims, kind), orc.build(prims, kind)
    same_rows(tc.spatial_crs(qb, PRED_BOX), to.spatial_crs(qb, PRED_BOX))
    qp = np.concatenate([clouds.uniform01(78, q // 2), prims[: q // 2, :3]]).astype(F)
    same_rows(tc.spatial_crs(qp, PRED_POINT), to.spatial_crs(qp, PRED_POINT))


def test_spatial_vs_bruteforce(cuda):
    n, q = 3000, 500
    pts = clouds.uniform01(1, n)
    spheres = np.concatenate([clouds.uniform01(2, q), np.full((q, 1), 0.1, F)], 1).astype(F)
    t = cuda.build(pts, PRIM_POINT)
    off, idx = t.spatial_crs(spheres, PRED_SPHERE)
    assert rows_of(off, idx) == brute.rows_from_mask(brute.spheres_vs_points(spheres, pts))
    boxes = boxes_from(3, n, 0.05)
    tb = cuda.build(boxes, PRIM_BOX)
    off, idx = tb.spatial_crs(spheres, PRED_SPHERE)
    assert rows_of(off, idx) == brute.rows_from_mask(brute.spheres_vs_boxes(spheres, boxes))
    qb = boxes_from(4, q, 0.1)
    off, idx = tb.spatial_crs(qb, PRED_BOX)
    assert rows_of(off, idx) == brute.rows_from_mask(brute.boxes_vs_boxes(qb, boxes))


def test_spatial_boundary_radius_exact(cuda, orc):
    # radii equal to exact neighbour distances: the <= comparison must agree bit for bit
    pts = clouds.uniform01(8, 20_000)
    to = orc.build(pts, PRIM_POINT)
    off, idx, dist = to.nearest_crs(pts[:5000], 6)
    radii = dist.reshape(-1, 6)[:, 5:6]
    spheres = np.concatenate([pts[:5000], radii], 1).astype(F)
    tc = cuda.build(pts, PRIM_POINT)
    same_rows(tc.spatial_crs(spheres, PRED_SPHERE), to.spatial_crs(spheres, PRED_SPHERE))


def test_bvh_driver_1m_spatial_counts(cuda, orc):
    # config 1 shape (BASELINE.json): 1M filled-box points, 100k of the queries checked
    n = 1_000_000
    pts = clouds.filled_box(0x5EED0001, n)
    qs = clouds.filled_box(0x5EED0002, 100_000)
    r = clouds.bvh_driver_radius(10)
    spheres = np.concatenate([qs, np.full((len(qs), 1), r, F)], 1).astype(F)
    tc, to = cuda.build(pts, PRIM_POINT), orc.build(pts, PRIM_POINT)
    same_rows(tc.spatial_crs(spheres, PRED_SPHERE), to.spatial_crs(spheres, PRED_SPHERE))


# ---------------------------------------------------------------- nearest ----
def check_knn_vs_oracle(res_c, res_o):
    (oc, ic, dc), (oo, io, do) = res_c, res_o
    assert np.array_equal(oc, oo)
    # distances are the same arithmetic on both sides: bit-exact
    assert np.array_equal(dc, do)
    # indices equal except inside groups of exactly tied distances
    diff = ic != io
    if diff.any():
        rows = np.repeat(np.arange(len(oc) - 1), np.diff(oc))
        for r in np.unique(rows[diff]):
            s, e = oc[r], oc[r + 1]
            dd = dc[s:e]
            tied = np.zeros(e - s, bool)
            tied[1:] |= dd[1:] == dd[:-1]
            tied[:-1] |= dd[1:] == dd[:-1]
            tied[-1] = True  # the k-th place may be tied with an excluded candidate
            assert np.all(tied[ic[s:e] != io[s:e]]), r


@pytest.mark.parametrize("name,kind,gen", PRIM_CASES, ids=[c[0] for c in PRIM_CASES])
@pytest.mark.parametrize("k", [1, 3, 10, 16, 32, 50])
def test_nearest_vs_oracle(cuda, orc, name, kind, gen, k):
    n, q = 40_000, 8000
    prims = gen(17, n)
    if kind == PRIM_POINT:
        prims = (prims / F(np.cbrt(n)) * F(0.5) + F(0.5)).astype(F)
    qp = (clouds.uniform01(55, q) * F(1.2) - F(0.1)).astype(F)
    tc, to = cuda.build(prims, kind), orc.build(prims, kind)
    check_knn_vs_oracle(tc.nearest_crs(qp, k), to.nearest_crs(qp, k))


def test_nearest_bruteforce_and_short_rows(cuda):
    n, q = 2000, 300
    pts = clouds.uniform01(31, n)
    qp = clouds.uniform01(32, q)
    t = cuda.build(pts, PRIM_POINT)
    D = brute.dist_point_point(qp, pts)
    for k in (1, 10, 40):
        off, idx, dist = t.nearest_crs(qp, k)
        brute.knn_check(D, k, off, idx, dist)
    # k > n: rows are shorter than k
    small = cuda.build(pts[:7], PRIM_POINT)
    off, idx, dist = small.nearest_crs(qp[:20], 10)
    assert list(off) == [7 * i for i in range(21)]
    brute.knn_check(D[:20, :7], 10, off, idx, dist)
    # per-query k, including k = 0
    ks = (np.arange(q) % 5).astype(np.int32) * 3
    off, idx, dist = t.nearest_crs(qp, ks)
    brute.knn_check(D, ks, off, idx, dist)


def test_nearest_duplicates(cuda):
    pts = np.repeat(clouds.uniform01(5, 100), 30, axis=0)
    t = cuda.build(pts, PRIM_POINT)
    qp = clouds.uniform01(6, 200)
    off, idx, dist = t.nearest_crs(qp, 10)
    brute.knn_check(brute.dist_point_point(qp, pts), 10, off, idx, dist)


# ---------------------------------------------------------------- half / dbscan ----
def test_half_traversal_vs_oracle(cuda, orc):
    pts = clouds.uniform01(12, 30_000)
    r = 0.02
    pc = cuda.build(pts, PRIM_POINT).half_pairs(r)
    po = orc.build(pts, PRIM_POINT).half_pairs(r)
    canon = lambda p: np.unique(np.sort(p.astype(np.int64), 1), axis=0)
    assert len(pc) == len(po)
    assert np.array_equal(canon(pc), canon(po))
    assert len(canon(pc)) == len(pc)  # each pair exactly once


@pytest.mark.parametrize("impl", [0, 1], ids=["fdbscan", "densebox"])
@pytest.mark.parametrize("minpts", [2, 5, 10])
@pytest.mark.parametrize("algo", [0, 1], ids=["dbscan", "dbscan_star"])
def test_dbscan_vs_oracle(cuda, impl, minpts, algo):
    n = 100_000
    pts = clouds.clustered(11, n, n_clusters=8, domain=1.0e4, spread=20.0, noise_frac=0.02)
    labels = cuda.dbscan(pts, 25.0, minpts, impl, algo)
    dbscan_checks.check_against_oracle(pts, 25.0, minpts, labels, impl, algo)


@pytest.mark.parametrize("impl", [0, 1], ids=["fdbscan", "densebox"])
def test_dbscan_uniform_sparse(cuda, impl):
    n = 200_000
    pts = clouds.filled_box(21, n)
    eps = float(clouds.bvh_driver_radius(10)) * 0.5
    for minpts in (2, 4):
        labels = cuda.dbscan(pts, eps, minpts, impl, 0)
        dbscan_checks.check_against_oracle(pts, eps, minpts, labels, impl, 0)


def test_dbscan_densebox_precision_guard(cuda):
    pts = (clouds.uniform01(1, 1000) + F(1.0e7)).astype(F)
    with pytest.raises(RuntimeError):
        cuda.dbscan(pts, 1e-3, 5, 1, 0)


# ---------------------------------------------------------------- host-buffer path ----
def test_host_buffer_entry_points(cuda, orc):
    import torch
    import arborx_b200 as abx
    n, q = 20_000, 5000
    pts = clouds.uniform01(41, n)
    spheres = np.concatenate([clouds.uniform01(42, q), np.full((q, 1), 0.05, F)], 1).astype(F)
    space = abx.ExecutionSpace()
    bvh = abx.BoundingVolumeHierarchy(space, torch.from_numpy(pts))  # host primitives
    idx, off = bvh.query(space, abx.intersects(torch.from_numpy(spheres)))
    assert not idx.is_cuda and not off.is_cuda
    to = orc.build(pts, PRIM_POINT)
    same_rows((off.numpy(), idx.numpy().view(np.uint32)), to.spatial_crs(spheres, PRED_SPHERE))
    idx, off, dist = bvh.query(space, abx.nearest(torch.from_numpy(spheres[:, :3].copy()), 5), return_distances=True)
    check_knn_vs_oracle((off.numpy(), idx.numpy().view(np.uint32), dist.numpy()), to.nearest_crs(spheres[:, :3], 5))
    labels = abx.dbscan(space, torch.from_numpy(pts), 0.02, 4, abx.DBSCANParameters(0, 0))
    dbscan_checks.check_against_oracle(pts, 0.02, 4, labels.numpy(), 0, 0, verify=False)


def test_nearest_short_rows_for_unreachable_leaves(cuda, orc):
    # an "empty" box (ArborX_Box.hpp:35-44) is at infinite distance from everything: it is never
    # accepted (distance < radius fails), so rows come out shorter than min(k, n) and are compacted
    # (DistributedTree's top tree holds such boxes for ranks without primitives)
    fm = np.finfo(F).max
    boxes = np.array([[0, 0, 0, 1, 1, 1], [fm, fm, fm, -fm, -fm, -fm], [2, 2, 2, 3, 3, 3],
                      [fm, fm, fm, -fm, -fm, -fm]], F)
    qp = np.array([[0.5, 0.5, 0.5], [2.5, 2.5, 2.9], [9, 9, 9]], F)
    tc, to = cuda.build(boxes, PRIM_BOX), orc.build(boxes, PRIM_BOX)
    for k in (1, 2, 3, 4, 40):
        oc, ic, dc = tc.nearest_crs(qp, k)
        oo, io, do = to.nearest_crs(qp, k)
        assert np.array_equal(oc, oo) and np.array_equal(ic, io) and np.array_equal(dc, do), k
    ks = np.array([4, 0, 2], np.int32)
    oc, ic, dc = tc.nearest_crs(qp, ks)
    oo, io, do = to.nearest_crs(qp, ks)
    assert np.array_equal(oc, oo) and np.array_equal(ic, io) and np.array_equal(dc, do)


def test_neighbor_lists(cuda):
    """findHalfNeighborList / findFullNeighborList (ArborX_NeighborList.hpp:47-190; tstNeighborList.cpp compares the
    full list with a radius query and the half list with its j < i filter): here against the O(n^2) matrix."""
    import torch
    import arborx_b200 as abx
    n, r = 3000, 0.08
    pts = clouds.uniform01(11, n)
    space = abx.ExecutionSpace()
    d = brute.dist_point_point(pts, pts) <= F(r)
    np.fill_diagonal(d, False)
    off, idx = abx.find_full_neighbor_list(space, torch.from_numpy(pts).cuda(), r)
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    for i in range(n):
        assert np.array_equal(np.sort(idx[off[i]:off[i + 1]]), np.nonzero(d[i])[0]), i
    hoff, hidx = abx.find_half_neighbor_list(space, torch.from_numpy(pts).cuda(), r)
    hoff, hidx = hoff.cpu().numpy(), hidx.cpu().numpy()
    assert hoff[-1] * 2 == off[-1]
    seen = set()
    for i in range(n):
        for j in hidx[hoff[i]:hoff[i + 1]]:
            assert d[i, j] and (j, i) not in seen and (i, j) not in seen
            seen.add((i, int(j)))
