"""Golden / known-answer vectors restated from the reference's own tests and
examples.  Paths are relative to /root/reference.  Pure data
and tiny generators -- nothing here reads /root/reference at run time.
"""
import itertools

import numpy as np

F = np.float32

# test/tstDetailsMortonCodes.cpp:25-28, :37-40
EXPAND_BITS_32 = [(0b110010011101, 0b000000001000000001001001000001)]
EXPAND_BITS_64 = [(0b11111111111111000001,
                   0b001001001001001001001001001001001001001001000000000000000001)]
# test/tstDetailsMortonCodes.cpp:73-76, :78-81 (points in the unit cube)
MORTON32 = [((0, 0, 0), 0), ((1, 1, 1), 0x3FFFFFFF), ((0, 0, 1), 0x9249249), ((1, 1, 0), 0x36DB6DB6)]
MORTON64 = [((0, 0, 0), 0), ((1, 1, 1), 0x7FFFFFFFFFFFFFFF), ((0, 0, 1), 0x1249249249249249),
            ((1, 1, 0), 0x6DB6DB6DB6DB6DB6)]

# test/tstDetailsTreeConstruction.cpp:33-102 (assign_morton_codes)
_N21 = float(1 << 21)
ASSIGN_MORTON_POINTS = np.array([[0.0, 0.0, 0.0], [0.25, 0.75, 0.25], [0.75, 0.25, 0.25], [0.75, 0.75, 0.25],
                                 [1.33, 2.33, 3.33], [1.66, 2.66, 3.66], [_N21, _N21, _N21]], F)
ASSIGN_MORTON_ANCHORS = [(0, 0, 0)] * 4 + [(1, 2, 3)] * 2 + [((1 << 21) - 1,) * 3]
ASSIGN_MORTON_SCENE = np.array([0, 0, 0, _N21, _N21, _N21], F)

# test/tstDetailsTreeConstruction.cpp:119-149 (indirect_sort)
INDIRECT_SORT = dict(keys=[4, 3, 2, 1], sorted=[1, 2, 3, 4], perm=[3, 2, 1, 0])

# test/tstDetailsTreeConstruction.cpp:217-269 (Karras example)
KARRAS_CODES = [int(s, 2) for s in ["00001", "00010", "00100", "00101", "10011", "11000", "11001", "11110"]]
KARRAS_DFS = "I0I3I1L0L1I2L2L3I4L4I5I6L5L6L7"


def rope_dfs_string(n, leaf_rope, left_child):
    """test/tstDetailsTreeConstruction.cpp:177-202: print L<i> and follow the rope
    for a leaf; print I<i> and descend into left_child for an internal node."""
    out = []
    node = n
    while node != -1:
        if node < n:
            out.append("L%d" % node)
            node = int(leaf_rope[node])
        else:
            out.append("I%d" % (node - n))
            node = int(left_child[node - n])
    return "".join(out)


# test/tstQueryTreeDegenerate.cpp:135-225 (single leaf [0,1]^3)
ONE_LEAF = dict(
    boxes=np.array([[0, 0, 0, 1, 1, 1]], F),
    spheres=np.array([[0, 0, 0, 1], [1, 1, 1, 3], [5, 5, 5, 2]], F),
    spheres_offsets=[0, 1, 2, 2], spheres_indices=[0, 0],
    nearest_pts=np.array([[0, 0, 0], [4, 5, 1]], F), nearest_k=[3, 1],
    nearest_offsets=[0, 1, 2], nearest_indices=[0, 0],
)
# :229-332 (two leaves, points (0,0,0) and (1,1,1))
TWO_LEAVES = dict(
    points=np.array([[0, 0, 0], [1, 1, 1]], F),
    boxes_q=np.array([[0, 0, 0, 1, 1, 1], [.5, .5, .5, 1.5, 1.5, 1.5]], F),
    boxes_rows=[[0, 1], [1]],
    nearest_pts=np.array([[0, 0, 0], [1, 0, 0]], F), nearest_k=[2, 4],
    nearest_rows=[[0, 1], [0, 1]],
)
# :335-369 (duplicated leaves)
DUPLICATES = dict(
    boxes=np.array([[0, 0, 0, 0, 0, 0]] + [[1, 1, 1, 1, 1, 1]] * 3, F),
    spheres=np.array([[0, 0, 0, 1], [1, 1, 1, 1], [.5, .5, .5, 1]], F),
    offsets=[0, 1, 4, 8], rows=[[0], [1, 2, 3], [0, 1, 2, 3]],
)
# :375-446 (not exactly degenerate: chain of 4096 boxes [i,i+1]^3)
CHAIN_N = 4096


def chain_boxes(n=CHAIN_N):
    i = np.arange(n, dtype=F)
    return np.stack([i, i, i, i + 1, i + 1, i + 1], 1).astype(F)


# test/tstQueryTreeTraversalPolicy.cpp:32-129 (buffer optimisation)
BUFFER_POLICY = dict(
    points=np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]], F),
    # queries: {}, [0,3]^3-ish box catching all four, {}  (empty boxes never match)
    boxes_q=np.array([[np.finfo(F).max] * 3 + [-np.finfo(F).max] * 3,
                      [0, 0, 0, 3, 3, 3],
                      [np.finfo(F).max] * 3 + [-np.finfo(F).max] * 3], F),
    offsets=[0, 0, 4, 4], rows=[[], [0, 1, 2, 3], []],
    ok_buffers=[0, -4, 5, -5, 1], throwing_buffers=[-1],
)
# :131-202 (unsorted predicates)
PREDICATE_SORTING = dict(
    points=np.array([[i, i, i] for i in range(4)], F),
    boxes_q=np.array([[2, 2, 2, 3, 3, 3], [0, 0, 0, 1, 1, 1]], F), rows=[[2, 3], [0, 1]],
    nearest_pts=np.array([[2.5, 2.5, 2.5], [0.5, 0.5, 0.5]], F), nearest_k=2, nearest_rows=[[2, 3], [0, 1]],
)

# test/tstDetailsHalfTraversal.cpp:60-106: 24 collinear points, every unordered pair once
HALF_TRAVERSAL_N = 24

# test/tstUnionFind.cpp:69-114
UNION_FIND = dict(n=5, steps=[
    ([(3, 0)], [0, 1, 2, 0, 4]),
    ([(1, 2), (4, 1)], [0, 1, 1, 0, 1]),
    ([(0, 1)], [0, 0, 0, 0, 0]),
])

# test/tstGeometryDistance.cpp:47-60 and tstGeometryIntersects.cpp:43-135 (subset)
DIST_POINT_BOX = dict(box=[-1, -1, -1, 1, 1, 1], cases=[
    ((0, 0, 0), 0.0), ((-1, -1, -1), 0.0), ((-2, -1, -1), 1.0), ((-2, -2, -1), np.sqrt(F(2))),
    ((-2, -2, -2), np.sqrt(F(3)))])
SPHERE_POINT = dict(sphere=[0, 0, 0, 1], hits=[(-.6, -.8, 0)], misses=[(-.7, -.8, 0)])

# examples/simple_intersection/example_intersection.cpp:30-81 is 2-D boxes; lifted to z=[0,0]
# test/tstQueryTreeManufacturedSolution.cpp:29-272


def structured_grid(nx=11, ny=11, nz=11, L=100.0):
    hx, hy, hz = F(L / (nx - 1)), F(L / (ny - 1)), F(L / (nz - 1))
    n = nx * ny * nz
    pts = np.zeros((n, 3), F)
    for i, j, k in itertools.product(range(nx), range(ny), range(nz)):
        pts[i + j * nx + k * nx * ny] = (F(i) * hx, F(j) * hy, F(k) * hz)
    return pts, (hx, hy, hz)


def structured_grid_neighbor_queries(nx=11, ny=11, nz=11, L=100.0):
    """(ii): boxes reaching the first neighbours and the expected 27-stencil rows."""
    pts, (hx, hy, hz) = structured_grid(nx, ny, nz, L)
    n = nx * ny * nz
    boxes = np.zeros((n, 6), F)
    rows = [None] * n
    for i, j, k in itertools.product(range(nx), range(ny), range(nz)):
        idx = i + j * nx + k * nx * ny
        boxes[idx] = (F(i - 1) * hx, F(j - 1) * hy, F(k - 1) * hz, F(i + 1) * hx, F(j + 1) * hy, F(k + 1) * hz)
        r = []
        for di, dj, dk in itertools.product((-1, 0, 1), repeat=3):
            a, b, c = i + di, j + dj, k + dk
            if 0 <= a < nx and 0 <= b < ny and 0 <= c < nz:
                r.append(a + b * nx + c * nx * ny)
        rows[idx] = sorted(r)
    return pts, boxes, rows


def structured_grid_random_boxes(seed=0, nx=11, ny=11, nz=11, L=100.0):
    """(iii): boxes of one grid step centred within 0.45 step of a lattice node."""
    pts, (hx, hy, hz) = structured_grid(nx, ny, nz, L)
    n = nx * ny * nz
    rng = np.random.default_rng(seed)
    ijk = np.stack([rng.integers(0, nx, n), rng.integers(0, ny, n), rng.integers(0, nz, n)], 1)
    shift = rng.uniform(-0.45, 0.45, (n, 3))
    c = ((ijk + shift) * np.array([hx, hy, hz])).astype(F)
    half = np.array([hx, hy, hz], F) / F(2)
    boxes = np.concatenate([c - half, c + half], 1).astype(F)
    expect = ijk[:, 0] + ijk[:, 1] * nx + ijk[:, 2] * nx * ny
    return pts, boxes, expect


# ---- DBSCAN ---------------------------------------------------------------
SQ3, SQ12, SQ48 = F(np.sqrt(3)), F(np.sqrt(12)), F(np.sqrt(48))
DB_P2 = np.array([[0, 0, 0], [1, 1, 1]], F)
DB_P4 = np.array([[0, 0, 0], [1, 1, 1], [3, 3, 3], [6, 6, 6]], F)
DB_BRIDGE = np.array([[-1, .5, 0], [-1, -.5, 0], [-1, 0, 0], [0, 0, 0], [1, 0, 0], [1, .5, 0], [1, -.5, 0]], F)
DB_STRIPPED = np.array([[0, -2, 0], [-1, -2, 0], [1, -2, 0], [0, -1, 0],
                        [0, 2, 0], [-1, 2, 0], [1, 2, 0], [0, 1, 0],
                        [2, 0, 0], [2, -1, 0], [2, 1, 0], [1, 0, 0],
                        [0, 0, 0]], F)
# test/tstDBSCAN.cpp:44-163: (points, eps, minpts, labels, algo, expected verifier verdict)
VERIFIER_CASES = [
    (DB_P2, SQ3 - F(0.1), 2, [-1, -1], "dbscan", True),
    (DB_P2, SQ3 - F(0.1), 2, [-1, -1], "dbscan*", True),
    (DB_P2, SQ3 - F(0.1), 2, [1, 2], "dbscan", False),
    (DB_P2, SQ3 - F(0.1), 2, [1, 1], "dbscan*", False),
    (DB_P2, SQ3, 2, [1, 1], "dbscan", True),
    (DB_P2, SQ3, 2, [1, 1], "dbscan*", True),
    (DB_P2, SQ3, 2, [1, 2], "dbscan", False),
    (DB_P2, SQ3, 2, [1, 2], "dbscan*", False),
    (DB_P2, SQ3, 3, [-1, -1], "dbscan", True),
    (DB_P2, SQ3, 3, [-1, -1], "dbscan*", True),
    (DB_P2, SQ3, 3, [1, 1], "dbscan", False),
    (DB_P2, SQ3, 3, [1, 1], "dbscan*", False),
    (DB_P4, SQ3, 2, [1, 1, -1, -1], "dbscan", True),
    (DB_P4, SQ3, 2, [1, 1, -1, -1], "dbscan*", True),
    (DB_P4, SQ3, 3, [-1, -1, -1, -1], "dbscan", True),
    (DB_P4, SQ3, 3, [-1, -1, -1, -1], "dbscan*", True),
    (DB_P4, SQ12, 2, [3, 3, 3, -1], "dbscan", True),
    (DB_P4, SQ12, 2, [3, 3, 3, -1], "dbscan*", True),
    (DB_P4, SQ12, 3, [3, 3, 3, -1], "dbscan", True),
    (DB_P4, SQ12, 3, [-1, 3, -1, -1], "dbscan*", True),
    (DB_P4, SQ12, 4, [-1, -1, -1, -1], "dbscan", True),
    (DB_P4, SQ12, 4, [-1, -1, -1, -1], "dbscan*", True),
    (DB_P4, SQ48, 2, [5, 5, 5, 5], "dbscan", True),
    (DB_P4, SQ48, 2, [5, 5, 5, 5], "dbscan*", True),
    (DB_P4, SQ48, 3, [5, 5, 5, 5], "dbscan", True),
    (DB_P4, SQ48, 3, [5, 5, 5, -1], "dbscan*", True),
    (DB_P4, SQ48, 4, [7, 7, 7, 7], "dbscan", True),
    (DB_P4, SQ48, 4, [-1, -1, 7, -1], "dbscan*", True),
    (DB_P4, SQ48, 5, [-1, -1, -1, -1], "dbscan", True),
    (DB_P4, SQ48, 5, [-1, -1, -1, -1], "dbscan*", True),
    (DB_BRIDGE, F(1), 3, [5, 5, 5, 5, 5, 5, 5], "dbscan", True),
    (DB_BRIDGE, F(1), 3, [5, 5, 5, 5, 5, 5, 5], "dbscan*", True),
    (DB_BRIDGE, F(1), 4, [5, 5, 5, 5, 6, 6, 6], "dbscan", True),
    (DB_BRIDGE, F(1), 4, [5, 5, 5, 6, 6, 6, 6], "dbscan", True),
    (DB_BRIDGE, F(1), 4, [-1, -1, 5, -1, 6, -1, -1], "dbscan*", True),
    (DB_STRIPPED, F(1), 4, [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 5], "dbscan", True),
    (DB_STRIPPED, F(1), 4, [0, -1, -1, -1, 1, -1, -1, -1, 2, -1, -1, -1, 5], "dbscan*", True),
    (DB_STRIPPED, F(1), 4, [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, -1], "dbscan", False),
    (DB_STRIPPED, F(1), 4, [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, -1], "dbscan*", False),
]
# test/tstDBSCAN.cpp:177-347: (points, eps, minpts) that must pass the verifier for
# both implementations x both algorithms
SQ31 = F(np.sqrt(3.1))
DBSCAN_RUN_CASES = [
    (DB_P2, SQ31 - F(0.1), 2), (DB_P2, SQ31, 2), (DB_P2, SQ31, 3),
    (DB_P4, SQ31, 2), (DB_P4, SQ31, 3), (DB_P4, F(2) * SQ31, 2), (DB_P4, F(2) * SQ31, 3), (DB_P4, F(2) * SQ31, 4),
    (DB_P4, F(3) * SQ31, 2), (DB_P4, F(3) * SQ31, 3), (DB_P4, F(3) * SQ31, 4), (DB_P4, F(3) * SQ31, 5),
    (DB_BRIDGE, F(1), 3), (DB_BRIDGE, F(1), 4), (DB_STRIPPED, F(1), 4),
]
# examples/dbscan/example_dbscan.cpp:33-90 (2-D points lifted to z = 0)
EXAMPLE_DBSCAN_POINTS = np.array([[4, 3, 0], [0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0], [3, 3, 0], [3, 4, 0],
                                  [4, 4, 0], [4, 0, 0], [2, 2, 0]], F)
EXAMPLE_DBSCAN = [
    (1.0, 2, [[0, 1, 1, 1, 1, 0, 0, 0, -1, -1]]),
    (1.0, 5, [[-1] * 10]),
    (1.5, 2, [[0, 0, 0, 0, 0, 0, 0, 0, -1, 0]]),
    (1.5, 4, [[0, 1, 1, 1, 1, 0, 0, 0, -1, 0], [0, 1, 1, 1, 1, 0, 0, 0, -1, 1]]),
]
# benchmarks/cluster/input.txt with --eps=1.4 --verify (benchmarks/cluster/CMakeLists.txt:13)
BENCH_INPUT_POINTS = np.array([[3, 2, 0], [0, 0, 0], [0, 1, 0], [1, 1, 0], [1, 0, 0], [2, 2, 0], [2, 3, 0],
                               [3, 3, 0]], F)
