"""Known-answer tests restated from the reference's own test-suite (SURVEY.md
App. B; data in tests/golden/vectors.py).  Each test runs against the CPU oracle
(-m "not gpu": this pins the oracle) and against the CUDA product through its
C ABI (-m gpu: parity on the reference's golden vectors)."""
import numpy as np
import pytest

from tests.engines import (PRED_BOX, PRED_POINT, PRED_SPHERE, PRIM_BOX, PRIM_POINT, SearchException, rows_of)
from tests.golden import vectors as V

F = np.float32


def test_assign_morton_codes(engine):
    # test/tstDetailsTreeConstruction.cpp:33-102 (boxes around points)
    boxes = np.concatenate([V.ASSIGN_MORTON_POINTS, V.ASSIGN_MORTON_POINTS], 1)
    scene = engine.scene_bounds(boxes, PRIM_BOX)
    assert np.array_equal(scene, V.ASSIGN_MORTON_SCENE)
    codes = engine.morton64_codes(boxes, scene, PRIM_BOX)
    import oracle
    e = oracle.lib().orc_expand_bits2_64
    ref = [4 * e(i) + 2 * e(j) + e(k) for (i, j, k) in V.ASSIGN_MORTON_ANCHORS]
    assert [int(c) for c in codes] == ref
    codes_p = engine.morton64_codes(V.ASSIGN_MORTON_POINTS, scene, PRIM_POINT)
    assert [int(c) for c in codes_p] == ref


def test_indirect_sort(engine):
    keys, perm = engine.sort_u64(np.array(V.INDIRECT_SORT["keys"], np.uint64))
    assert list(keys) == V.INDIRECT_SORT["sorted"]
    assert list(perm) == V.INDIRECT_SORT["perm"]


def test_karras_example(engine):
    n = len(V.KARRAS_CODES)
    prims = np.zeros((n, 3), F)
    t = engine.from_sorted_codes(prims, np.array(V.KARRAS_CODES, np.uint64))
    d = t.export()
    assert V.rope_dfs_string(n, d["leaf_rope"], d["left_child"]) == V.KARRAS_DFS


def test_empty_tree(engine):
    # test/tstQueryTreeDegenerate.cpp:25-133
    t = engine.build(np.zeros((0, 3), F), PRIM_POINT)
    assert t.n == 0
    off, idx = t.spatial_crs(np.array([[0, 0, 0, 1], [1, 1, 1, 2]], F), PRED_SPHERE)
    assert list(off) == [0, 0, 0] and len(idx) == 0
    off, idx, dist = t.nearest_crs(np.array([[0, 0, 0], [1, 1, 1]], F), 3)
    assert list(off) == [0, 0, 0] and len(idx) == 0
    off, idx = t.spatial_crs(np.zeros((0, 4), F), PRED_SPHERE)
    assert list(off) == [0]


def test_single_leaf(engine):
    g = V.ONE_LEAF
    t = engine.build(g["boxes"], PRIM_BOX)
    assert np.array_equal(t.bounds(), np.array([0, 0, 0, 1, 1, 1], F))
    off, idx = t.spatial_crs(g["spheres"], PRED_SPHERE)
    assert list(off) == g["spheres_offsets"] and list(idx) == g["spheres_indices"]
    off, idx, _ = t.nearest_crs(g["nearest_pts"], np.array(g["nearest_k"], np.int32))
    assert list(off) == g["nearest_offsets"] and list(idx) == g["nearest_indices"]
    off, idx = t.spatial_crs(np.zeros((0, 4), F), PRED_SPHERE)
    assert list(off) == [0]


def test_two_leaves(engine):
    g = V.TWO_LEAVES
    for kind, prims in ((PRIM_POINT, g["points"]), (PRIM_BOX, np.concatenate([g["points"]] * 2, 1))):
        t = engine.build(prims, kind)
        assert np.array_equal(t.bounds(), np.array([0, 0, 0, 1, 1, 1], F))
        off, idx = t.spatial_crs(g["boxes_q"], PRED_BOX)
        assert rows_of(off, idx) == g["boxes_rows"]
        off, idx, _ = t.nearest_crs(g["nearest_pts"], np.array(g["nearest_k"], np.int32))
        assert rows_of(off, idx) == g["nearest_rows"]


def test_duplicated_leaves(engine):
    g = V.DUPLICATES
    t = engine.build(g["boxes"], PRIM_BOX)
    off, idx = t.spatial_crs(g["spheres"], PRED_SPHERE)
    assert list(off) == g["offsets"]
    assert rows_of(off, idx) == g["rows"]


def test_chain_not_degenerate(engine):
    n = V.CHAIN_N
    t = engine.build(V.chain_boxes(n), PRIM_BOX)
    off, idx = t.spatial_crs(np.array([[0, 0, 0, n, n, n]], F), PRED_BOX)
    assert rows_of(off, idx) == [list(range(n))]
    off, idx, dist = t.nearest_crs(np.array([[0, 0, 0]], F), n)
    assert rows_of(off, idx) == [list(range(n))]
    assert np.all(np.diff(dist) >= 0)


@pytest.mark.parametrize("sort_predicates", [True, False])
def test_structured_grid(engine, sort_predicates):
    pts, _ = V.structured_grid()
    n = pts.shape[0]
    boxes = np.concatenate([pts, pts], 1)
    t = engine.build(boxes, PRIM_BOX)
    # (i) self queries
    off, idx = t.spatial_crs(boxes, PRED_BOX, sort_predicates)
    assert list(off) == list(range(n + 1)) and list(idx) == list(range(n))
    # (ii) first neighbours
    _, qb, rows = V.structured_grid_neighbor_queries()
    off, idx = t.spatial_crs(qb, PRED_BOX, sort_predicates)
    assert rows_of(off, idx) == rows
    # (iii) random boxes each containing exactly one lattice node
    _, qb, expect = V.structured_grid_random_boxes()
    off, idx = t.spatial_crs(qb, PRED_BOX, sort_predicates)
    assert list(off) == list(range(n + 1)) and list(idx) == list(expect)
    # same tree over points
    tp = engine.build(pts, PRIM_POINT)
    off, idx = tp.spatial_crs(qb, PRED_BOX, sort_predicates)
    assert list(idx) == list(expect)


def test_buffer_policy(engine):
    g = V.BUFFER_POLICY
    for kind, prims in ((PRIM_POINT, g["points"]), (PRIM_BOX, np.concatenate([g["points"]] * 2, 1))):
        t = engine.build(prims, kind)
        for b in g["ok_buffers"]:
            off, idx = t.spatial_crs(g["boxes_q"], PRED_BOX, True, b)
            assert list(off) == g["offsets"], b
            assert rows_of(off, idx) == g["rows"], b
        for b in g["throwing_buffers"]:
            with pytest.raises(SearchException):
                t.spatial_crs(g["boxes_q"], PRED_BOX, True, b)


@pytest.mark.parametrize("sort_predicates", [True, False])
def test_unsorted_predicates(engine, sort_predicates):
    g = V.PREDICATE_SORTING
    t = engine.build(g["points"], PRIM_POINT)
    off, idx = t.spatial_crs(g["boxes_q"], PRED_BOX, sort_predicates)
    assert rows_of(off, idx) == g["rows"]
    off, idx, _ = t.nearest_crs(g["nearest_pts"], g["nearest_k"], sort_predicates)
    assert rows_of(off, idx) == g["nearest_rows"]


def test_half_traversal(engine):
    n = V.HALF_TRAVERSAL_N
    pts = np.array([[i, i, i] for i in range(n)], F)
    t = engine.build(pts, PRIM_POINT)
    pairs = t.half_pairs(1e6)
    got = sorted((min(a, b), max(a, b)) for a, b in pairs.tolist())
    assert got == [(i, j) for i in range(n) for j in range(i + 1, n)]


def test_geometry_known_answers():
    # oracle-only arithmetic pins (test/tstGeometryDistance.cpp:47-60,
    # tstGeometryIntersects.cpp:92-103); the CUDA path is pinned through queries
    import ctypes as C
    import oracle
    L = oracle.lib()

    def fp(a):
        a = np.ascontiguousarray(a, F)
        return a, a.ctypes.data_as(C.POINTER(C.c_float))

    box, bp = fp(V.DIST_POINT_BOX["box"])
    for p, d in V.DIST_POINT_BOX["cases"]:
        a, ap = fp(p)
        assert L.orc_distance_point_box(ap, bp) == F(d)
    s, sp = fp(V.SPHERE_POINT["sphere"])
    for p in V.SPHERE_POINT["hits"]:
        a, ap = fp(p)
        assert L.orc_intersects(PRED_SPHERE, sp, PRIM_POINT, ap) == 1
    for p in V.SPHERE_POINT["misses"]:
        a, ap = fp(p)
        assert L.orc_intersects(PRED_SPHERE, sp, PRIM_POINT, ap) == 0
    # point-triangle zones (test/tstGeometryDistance.cpp:120-175 style): unit right triangle in z=0
    tri, tp = fp([0, 0, 0, 1, 0, 0, 0, 1, 0])
    for p, d in [((0.25, 0.25, 0), 0.0), ((0.25, 0.25, 2), 2.0), ((-1, -1, 0), np.sqrt(F(2))), ((2, 0, 0), 1.0),
                 ((0, 2, 0), 1.0), ((0.5, -1, 0), 1.0), ((-1, 0.5, 0), 1.0), ((1, 1, 0), np.sqrt(F(0.5)))]:
        a, ap = fp(p)
        assert abs(L.orc_distance_point_triangle(ap, tp) - d) < 1e-6


def test_morton_bits():
    import oracle
    L = oracle.lib()
    for x, y in V.EXPAND_BITS_32:
        assert L.orc_expand_bits2_32(x) == y
    for x, y in V.EXPAND_BITS_64:
        assert L.orc_expand_bits2_64(x) == y
    for p, c in V.MORTON32:
        assert L.orc_morton32(*map(float, p)) == c
    for p, c in V.MORTON64:
        assert L.orc_morton64(*map(float, p)) == c


def test_union_find():
    import ctypes as C
    import oracle
    L = oracle.lib()
    labels = np.arange(V.UNION_FIND["n"], dtype=np.int32)
    lp = labels.ctypes.data_as(C.POINTER(C.c_int))
    for merges, expect in V.UNION_FIND["steps"]:
        for i, j in merges:
            L.orc_union_find_merge(lp, i, j)
        reps = [L.orc_union_find_representative(lp, i) for i in range(len(labels))]
        assert reps == expect
