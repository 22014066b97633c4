"""Size-independent properties at BASELINE.json's full size (10M points / queries): the oracle cannot
check every row at this size in test time, so the CUDA path is checked through invariants -- sortedness,
permutation, box containment, count = fill, self-hit, symmetry, kNN/radius cross-consistency, agreement of
the two DBSCAN implementations -- plus the oracle on a random sample of the same queries."""
import numpy as np
import pytest
import torch

from tests import clouds

pytestmark = pytest.mark.gpu
N = 10_000_000


@pytest.fixture(scope="module")
def setup():
    import arborx_b200 as abx
    space = abx.ExecutionSpace()
    pts_h = clouds.filled_box(0x5EED0001, N)
    pts = torch.from_numpy(pts_h).cuda()
    bvh = abx.BoundingVolumeHierarchy(space, pts)
    return abx, space, pts_h, pts, bvh


def test_tree_invariants_10m(setup):
    abx, space, pts_h, pts, bvh = setup
    d = bvh.export_reference_layout(space)
    n = N
    codes = d["codes"]
    assert bool((codes[1:] >= codes[:-1]).all())  # sorted (non-negative int64 view of 63-bit codes)
    perm = d["leaf_index"].long()
    assert torch.equal(torch.sort(perm).values, torch.arange(n, device="cuda"))
    boxes, lc = d["boxes"], d["left_child"].long()
    # root box = scene bounds = min/max of the points
    assert torch.equal(boxes[0, :3], pts.min(0).values) and torch.equal(boxes[0, 3:], pts.max(0).values)
    assert torch.equal(bvh.bounds().cuda(), boxes[0])
    # every internal node contains its left child (leaf point or internal box) ...
    leaf_pts = pts[perm]
    is_leaf = lc < n
    child_lo = torch.where(is_leaf.unsqueeze(1), leaf_pts[lc.clamp(max=n - 1)], boxes[(lc - n).clamp(min=0), :3])
    child_hi = torch.where(is_leaf.unsqueeze(1), leaf_pts[lc.clamp(max=n - 1)], boxes[(lc - n).clamp(min=0), 3:])
    assert bool((boxes[:, :3] <= child_lo).all()) and bool((boxes[:, 3:] >= child_hi).all())
    # ... and its right child = rope of the left child
    rope_leaf, rope_int = d["leaf_rope"].long(), d["rope"].long()
    rc = torch.where(is_leaf, rope_leaf[lc.clamp(max=n - 1)], rope_int[(lc - n).clamp(min=0)])
    assert bool((rc >= 0).all())
    r_leaf = rc < n
    rlo = torch.where(r_leaf.unsqueeze(1), leaf_pts[rc.clamp(max=n - 1)], boxes[(rc - n).clamp(min=0), :3])
    rhi = torch.where(r_leaf.unsqueeze(1), leaf_pts[rc.clamp(max=n - 1)], boxes[(rc - n).clamp(min=0), 3:])
    assert bool((boxes[:, :3] <= rlo).all()) and bool((boxes[:, 3:] >= rhi).all())
    # the box is tight: it equals the union of its two children
    assert torch.equal(boxes[:, :3], torch.minimum(child_lo, rlo)) and torch.equal(boxes[:, 3:], torch.maximum(child_hi, rhi))
    # every node except the root is the child of exactly one node
    seen = torch.zeros(2 * n - 1, dtype=torch.int32, device="cuda")
    seen.index_add_(0, lc, torch.ones_like(lc, dtype=torch.int32))
    seen.index_add_(0, rc, torch.ones_like(rc, dtype=torch.int32))
    assert int(seen[n]) == 0 and bool((seen[:n] == 1).all()) and bool((seen[n + 1:] == 1).all())


def test_radius_properties_10m(setup):
    abx, space, pts_h, pts, bvh = setup
    r = float(clouds.bvh_driver_radius(10))
    preds = abx.make_intersects(pts, r)  # queries = values: each row contains its own index, relation symmetric
    idx, off = bvh.query(space, preds)
    counts = bvh.count(space, preds)
    off64 = off.long()
    assert int(off[0]) == 0 and torch.equal(off64[1:] - off64[:-1], counts.long())
    assert int(off[-1]) == idx.numel()
    rows = torch.repeat_interleave(torch.arange(N, device="cuda"), counts.long())
    cols = idx.long()
    assert int((rows == cols).sum()) == N  # self hit, exactly once per row
    # symmetry: the multiset of (i, j) equals the multiset of (j, i)
    a = torch.sort(rows * N + cols).values
    b = torch.sort(cols * N + rows).values
    assert torch.equal(a, b)
    assert bool((a[1:] != a[:-1]).all())  # no duplicates inside a row
    # unsorted predicates and the count-up-to-N form agree
    idx2, off2 = bvh.query(space, preds, abx.TraversalPolicy(0, False))
    assert torch.equal(off2, off)
    lim = bvh.count(space, preds, limit=3)
    assert torch.equal(lim, torch.clamp(counts, max=3))


def test_radius_oracle_sample_10m(setup):
    """The bench's own radius query (10M spheres of the second cloud, r = cbrt(10 * 6 / pi)) against the oracle on a
    random sample of 20 000 of those queries: identical index sets per query."""
    import oracle
    abx, space, pts_h, pts, bvh = setup
    r = clouds.bvh_driver_radius(10)
    qs_h = clouds.filled_box(0x5EED0002, N)
    spheres_h = np.concatenate([qs_h, np.full((N, 1), r, np.float32)], 1).astype(np.float32)
    idx, off = bvh.query(space, abx.intersects(torch.from_numpy(spheres_h).cuda()))
    sel = np.sort(np.random.default_rng(1).choice(N, 20_000, replace=False))
    roff, ridx = oracle.Tree(pts_h).spatial_crs(spheres_h[sel])
    off_h = off.cpu().numpy().astype(np.int64)
    assert np.array_equal(np.diff(roff), (off_h[sel + 1] - off_h[sel]))
    # gather the sampled rows on the device, compare as sorted rows
    starts = torch.from_numpy(off_h[sel]).cuda()
    lens = torch.from_numpy(np.diff(roff).astype(np.int64)).cuda()
    pos = torch.repeat_interleave(starts - torch.cumsum(lens, 0) + lens, lens) + torch.arange(int(lens.sum()), device="cuda")
    got = idx[pos].cpu().numpy().view(np.uint32)
    row = np.repeat(np.arange(len(sel)), np.diff(roff))
    assert np.array_equal(got[np.lexsort((got, row))], ridx[np.lexsort((ridx, row))])


def test_knn_properties_10m(setup):
    abx, space, pts_h, pts, bvh = setup
    k = 10
    qs_h = clouds.filled_box(0x5EED0002, N)
    qs = torch.from_numpy(qs_h).cuda()
    idx, off, dist = bvh.query(space, abx.make_nearest(qs, k), return_distances=True)
    assert torch.equal(off.long(), torch.arange(N + 1, device="cuda") * k)
    d = dist.view(N, k)
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    # reported distances are the distances to the reported points (same operation order)
    p = pts[idx.long()].view(N, k, 3)
    t = p - qs.unsqueeze(1)
    d2 = t[..., 0] * t[..., 0]
    d2 = d2 + t[..., 1] * t[..., 1]
    d2 = d2 + t[..., 2] * t[..., 2]
    assert torch.equal(torch.sqrt(d2), d)
    # cross-check with the radius query: exactly >= k points lie within the k-th distance, and fewer than k
    # strictly inside the (k-1)-th ... checked on a 1M slice to bound memory
    m = 1_000_000
    spheres = torch.cat([qs[:m], d[:m, k - 1:k]], 1).contiguous()
    cnt = bvh.count(space, abx.intersects(spheres))
    assert bool((cnt >= k).all())
    # oracle on a random sample of the same queries (bit-exact distances)
    import oracle
    sel = np.random.default_rng(0).choice(N, 20_000, replace=False)
    ref = oracle.Tree(pts_h)
    roff, ridx, rd = ref.nearest_crs(qs_h[sel], k)
    assert np.array_equal(rd.reshape(-1, k), d[torch.from_numpy(sel).cuda()].cpu().numpy())
    # queries = values: nearest neighbour of a point is itself at distance 0
    idx0, off0, dist0 = bvh.query(space, abx.make_nearest(pts[:m], 1), return_distances=True)
    assert bool((dist0 == 0).all())


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("minpts", [2, 5])
def test_dbscan_gantao_vs_oracle(impl, minpts):
    """ArborX::dbscan on the benchmark's clustered cloud (GanTao, eps = 200) against the oracle at 2M points: the
    same noise set, the same labels on core points (smallest index of the component), border points accepted by
    the reference's verifier."""
    import arborx_b200 as abx
    import oracle
    space = abx.ExecutionSpace()
    n, eps = 2_000_000, 200.0
    x_h = clouds.gan_tao(3, n)
    lab = abx.dbscan(space, torch.from_numpy(x_h).cuda(), eps, minpts, abx.DBSCANParameters(impl, 0)).cpu().numpy()
    ref, core = oracle.dbscan(x_h, eps, minpts, impl, 0, return_core=True)
    assert np.array_equal(lab == -1, ref == -1)
    assert np.array_equal(lab[core], ref[core])
    assert int((lab >= 0).sum()) > n // 2 and len(np.unique(lab[lab >= 0])) == len(np.unique(ref[ref >= 0]))
    # border points may join any adjacent cluster: checked by the verifier on a 200k-point sub-cloud (it is O(n * m))
    m = 200_000
    sub = abx.dbscan(space, torch.from_numpy(x_h[:m].copy()).cuda(), eps, minpts, abx.DBSCANParameters(impl, 0))
    assert oracle.dbscan_verify(x_h[:m], eps, minpts, sub.cpu().numpy(), 0) == 0


def test_dbscan_implementations_agree_10m():
    import arborx_b200 as abx
    space = abx.ExecutionSpace()
    n = 4_000_000
    x = torch.from_numpy(clouds.gan_tao(5, n)).cuda()
    la = abx.dbscan(space, x, 200.0, 2, abx.DBSCANParameters(0, 0))
    lb = abx.dbscan(space, x, 200.0, 2, abx.DBSCANParameters(1, 0))
    assert torch.equal(la, lb)  # minpts = 2: every non-noise point is core, labels = min index of the component
    assert int((la >= 0).sum()) > n // 2
    lc = abx.dbscan(space, x, 200.0, 5, abx.DBSCANParameters(0, 1))
    ld = abx.dbscan(space, x, 200.0, 5, abx.DBSCANParameters(1, 1))
    assert torch.equal(lc, ld)  # DBSCAN*: only core points are labelled, deterministically


def test_mst_vs_oracle_2m():
    """MinimumSpanningTree at 2M points: edge for edge and weight for weight against the oracle (the tree is unique
    under the reference's edge order, detail/ArborX_BoruvkaHelpers.hpp:37-105)."""
    import arborx_b200 as abx
    import oracle
    space = abx.ExecutionSpace()
    pts_h = clouds.gan_tao(3, 2_000_000)
    for k in (1, 4):
        mst = abx.MinimumSpanningTree(space, torch.from_numpy(pts_h).cuda(), k)
        space.fence()
        e, w = oracle.mst(pts_h, k)
        got = mst.edges.cpu().numpy().astype(np.int64)
        key_got = np.sort(np.minimum(got[:, 0], got[:, 1]) * (1 << 32) + np.maximum(got[:, 0], got[:, 1]))
        ee = e.astype(np.int64)
        key_ref = np.sort(np.minimum(ee[:, 0], ee[:, 1]) * (1 << 32) + np.maximum(ee[:, 0], ee[:, 1]))
        assert np.array_equal(key_got, key_ref)
        assert np.array_equal(np.sort(mst.weights.cpu().numpy()), np.sort(w))


def test_mst_properties_10m(setup):
    """10M points: n - 1 edges that connect everything, every weight is the distance of its end points, and no
    point has a neighbour closer than its lightest tree edge (the cut property at single vertices)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    abx, space, pts_h, pts, bvh = setup
    mst = abx.MinimumSpanningTree(space, pts)
    space.fence()
    e = mst.edges.cpu().numpy()
    w = mst.weights.cpu().numpy()
    assert e.shape == (N - 1, 2) and e.min() >= 0 and e.max() < N
    g = coo_matrix((np.ones(N - 1, np.int8), (e[:, 0], e[:, 1])), shape=(N, N))
    ncomp, _ = connected_components(g, directed=False)
    assert ncomp == 1
    diff = pts_h[e[:, 0]] - pts_h[e[:, 1]]
    d = np.sqrt((diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]).astype(np.float32)
    assert np.array_equal(d, w)
    # nearest other point of every vertex = its lightest incident tree edge
    lightest = np.full(N, np.inf, np.float32)
    np.minimum.at(lightest, e[:, 0], w)
    np.minimum.at(lightest, e[:, 1], w)
    idx, off, dist = bvh.query(space, abx.nearest(pts, 2), return_distances=True)
    nn = dist.view(-1, 2)[:, 1].cpu().numpy()
    assert np.array_equal(nn, lightest)
