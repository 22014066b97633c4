"""The DistributedTree exchange kernels one by one, single process, through the C ABI, against numpy
restatements (distributed/detail/ArborX_DistributedTreeUtils.hpp:229-342, ArborX_DistributedTreeNearest.hpp:131-233):
routing against R = 4 rank boxes (count + fill, spheres / boxes / points / points with a radius array),
the CRS merges (remote rows as CRS and as query-id-sorted records), the kNN candidate merge, local kNN rows
in (index, rank) form with padding, pair_with_rank."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import clouds

pytestmark = pytest.mark.gpu
F = np.float32


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@pytest.fixture(scope="module")
def ctx():
    import arborx_b200 as abx
    from arborx_b200 import _lib
    return abx, _lib, _lib.lib(), abx.ExecutionSpace()


RANK_BOXES = np.array([[0, 0, 0, 1, 1, 1],
                       [1, 0, 0, 2, 1, 1],
                       [3.0e38, 3.0e38, 3.0e38, -3.0e38, -3.0e38, -3.0e38],  # a rank without primitives
                       [0.5, 0.5, 1, 1.5, 1.5, 2]], F)


def _route_reference(kind, preds, boxes, self_rank, radius=None):
    """hits[i, r]: predicate i must be forwarded to rank r (float32 arithmetic of routeKernel)."""
    q, R = preds.shape[0], boxes.shape[0]
    hits = np.zeros((q, R), bool)
    for r in range(R):
        lo, hi = boxes[r, :3], boxes[r, 3:]
        if r == self_rank or (lo > hi).any():
            continue
        if kind == 1:
            hits[:, r] = ~((preds[:, :3] > hi) | (preds[:, 3:6] < lo)).any(1)
            continue
        c = preds[:, :3]
        p = np.minimum(np.maximum(c, lo), hi) - c
        d2 = np.zeros(q, F)
        for d in range(3):
            d2 = (d2 + p[:, d] * p[:, d]).astype(F)
        if kind == 0:
            rr = (radius if radius is not None else preds[:, 3]).astype(F)
            with np.errstate(over="ignore"):
                r2 = ((rr * rr).astype(F) * F(1.0001)).astype(F) + F(1e-30)
            hits[:, r] = (d2 <= r2) | ~(r2 < np.inf)
        else:
            hits[:, r] = d2 == 0
    return hits


@pytest.mark.parametrize("case", ["sphere", "box", "point", "radius_array"])
@pytest.mark.parametrize("self_rank", [0, 3])
def test_route_count_and_fill(ctx, case, self_rank):
    abx, _lib, L, space = ctx
    q = 20_000
    c = (clouds.uniform01(11, q) * F(2.4) - F(0.2)).astype(F)
    radius_t, stride = None, 0
    if case == "sphere":
        kind = 0
        rad = (clouds.uniform01(12, q, 1)[:, 0] * F(0.3)).astype(F)
        rad[::97] = np.inf
        preds = np.concatenate([c, rad[:, None]], 1).astype(F)
        ref = _route_reference(0, preds, RANK_BOXES, self_rank)
    elif case == "box":
        kind = 1
        ext = (clouds.uniform01(13, q) * F(0.2)).astype(F)
        preds = np.concatenate([c, c + ext], 1).astype(F)
        ref = _route_reference(1, preds, RANK_BOXES, self_rank)
    elif case == "point":
        kind = 2
        preds = c.copy()
        preds[::5] = np.round(preds[::5] * 2) / 2  # points exactly on box faces
        ref = _route_reference(2, preds, RANK_BOXES, self_rank)
    else:
        kind, stride = 0, 3  # points + strided radius array (the k-th distances of kNN rows, k = 3)
        preds = c.copy()
        rows = (clouds.uniform01(14, q) * F(0.25)).astype(F)
        rows[::89, 2] = np.inf
        radius_t = torch.from_numpy(rows.reshape(-1)).cuda()
        ref = _route_reference(0, preds, RANK_BOXES, self_rank, rows[:, 2])
    R = RANK_BOXES.shape[0]
    d_preds = torch.from_numpy(preds).cuda()
    d_boxes = torch.from_numpy(RANK_BOXES).cuda()
    counts = torch.empty(R, dtype=torch.int32, device="cuda")
    rp = C.c_void_p(radius_t.data_ptr() + 8) if radius_t is not None else None  # element [i * 3 + 2]
    _lib.check(L.abx_dist_route_count(space.handle, kind, _ptr(d_preds), q, rp, stride, _ptr(d_boxes), R, self_rank,
                                      _ptr(counts)))
    h_counts = counts.cpu().numpy()
    assert np.array_equal(h_counts, ref.sum(0))
    assert h_counts[self_rank] == 0 and h_counts[2] == 0 and h_counts.sum() > 0
    base = np.concatenate([[0], np.cumsum(h_counts)]).astype(np.int32)
    d_base = torch.from_numpy(base[:R].copy()).cuda()
    cursors = torch.empty(R, dtype=torch.int32, device="cuda")
    qids = torch.full((int(base[-1]),), -1, dtype=torch.int32, device="cuda")
    _lib.check(L.abx_dist_route_fill(space.handle, kind, _ptr(d_preds), q, rp, stride, _ptr(d_boxes), R, self_rank,
                                     _ptr(d_base), _ptr(cursors), _ptr(qids)))
    h = qids.cpu().numpy()
    for r in range(R):  # grouped by destination; the order inside a group is unspecified
        assert np.array_equal(np.sort(h[base[r]:base[r + 1]]), np.nonzero(ref[:, r])[0])


def _random_crs(rng, q, max_row, n_values):
    counts = rng.integers(0, max_row + 1, q)
    counts[rng.random(q) < 0.3] = 0
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return off, rng.integers(0, n_values, off[-1]).astype(np.int32)


def _merge_reference(q, loff, lidx, rank, roff, rvals):
    off = np.concatenate([[0], np.cumsum(np.diff(loff) + np.diff(roff))]).astype(np.int32)
    vals = np.zeros((off[-1], 2), np.int32)
    for i in range(q):
        nl = loff[i + 1] - loff[i]
        vals[off[i]:off[i] + nl, 0] = lidx[loff[i]:loff[i + 1]]
        vals[off[i]:off[i] + nl, 1] = rank
        vals[off[i] + nl:off[i + 1]] = rvals[roff[i]:roff[i + 1]]
    return off, vals


@pytest.mark.parametrize("q", [1, 31, 5000])
def test_merge_crs_and_merge_sorted(ctx, q):
    abx, _lib, L, space = ctx
    rng = np.random.default_rng(q)
    rank = 2
    loff, lidx = _random_crs(rng, q, 40, 1000)
    roff, ridx = _random_crs(rng, q, 6, 1000)
    rvals = np.stack([ridx, rng.choice([0, 1, 3], ridx.shape[0]).astype(np.int32)], 1)
    ref_off, ref_vals = _merge_reference(q, loff, lidx, rank, roff, rvals)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_loff, d_lidx, d_roff, d_rvals = cu(loff), cu(lidx), cu(roff), cu(rvals)
    for form in ("crs", "sorted"):
        out_off = torch.empty(q + 1, dtype=torch.int32, device="cuda")
        out_vals = torch.full((max(int(ref_off[-1]), 1), 2), -7, dtype=torch.int32, device="cuda")
        if form == "crs":
            _lib.check(L.abx_dist_merge_crs(space.handle, q, _ptr(d_loff), _ptr(d_lidx), rank, _ptr(d_roff),
                                            _ptr(d_rvals), _ptr(out_off), _ptr(out_vals)))
        else:
            ids = cu(np.repeat(np.arange(q), np.diff(roff)).astype(np.int32))  # records in query-id order
            _lib.check(L.abx_dist_merge_sorted(space.handle, q, _ptr(d_loff), _ptr(d_lidx), rank, ids.shape[0],
                                               _ptr(ids), _ptr(d_rvals), _ptr(out_off), _ptr(out_vals)))
        assert np.array_equal(out_off.cpu().numpy(), ref_off)
        assert np.array_equal(out_vals.cpu().numpy()[:ref_off[-1]], ref_vals)


@pytest.mark.parametrize("k", [1, 3, 10])
def test_knn_merge(ctx, k):
    abx, _lib, L, space = ctx
    rng = np.random.default_rng(100 + k)
    q = 4000
    dist = np.sort(rng.random((q, k)).astype(F), 1)
    vals = np.stack([rng.integers(0, 1 << 20, (q, k)), np.full((q, k), 1)], 2).astype(np.int32)
    short = rng.random(q) < 0.1  # padded rows (fewer than k local entries)
    for i in np.nonzero(short)[0]:
        keep = rng.integers(0, k)
        dist[i, keep:] = np.inf
        vals[i, keep:] = -1
    m_per = rng.integers(0, 7, q)
    m_per[rng.random(q) < 0.6] = 0
    ids = np.repeat(np.arange(q), m_per).astype(np.int32)
    cd = rng.random(ids.shape[0]).astype(F)
    tie = rng.random(ids.shape[0]) < 0.2  # exact ties with a local entry: the local one stays ahead
    cd[tie] = dist[ids[tie], 0]
    cv = np.stack([rng.integers(0, 1 << 20, ids.shape[0]), rng.choice([0, 2, 3], ids.shape[0])], 1).astype(np.int32)
    ref_d, ref_v = dist.copy(), vals.copy()
    start = np.concatenate([[0], np.cumsum(m_per)])
    for i in np.nonzero(m_per)[0]:
        d_all = np.concatenate([dist[i], cd[start[i]:start[i + 1]]])
        v_all = np.concatenate([vals[i], cv[start[i]:start[i + 1]]])
        order = np.argsort(d_all, kind="stable")[:k]  # stable: local entries first among equals, then arrival order
        ref_d[i], ref_v[i] = d_all[order], v_all[order]
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_vals, d_dist = cu(vals.reshape(-1, 2)), cu(dist.reshape(-1))
    d_ids, d_cv, d_cd = cu(ids), cu(cv), cu(cd)  # named: the buffers must outlive the launch
    _lib.check(L.abx_dist_knn_merge(space.handle, ids.shape[0], _ptr(d_ids), _ptr(d_cv), _ptr(d_cd), k,
                                    _ptr(d_vals), _ptr(d_dist)))
    assert np.array_equal(d_dist.cpu().numpy().reshape(q, k), ref_d)
    got_v = d_vals.cpu().numpy().reshape(q, k, 2)
    # equal distances among the candidates themselves may be ordered either way: compare as sets per distance
    same = (got_v == ref_v).all(2)
    for i, j in zip(*np.nonzero(~same)):
        grp = ref_d[i] == ref_d[i, j]
        assert sorted(map(tuple, got_v[i][grp])) == sorted(map(tuple, ref_v[i][grp]))


@pytest.mark.parametrize("n,k", [(5000, 4), (3, 5), (1, 2), (0, 3)])
def test_nearest_pairs_rows(ctx, n, k):
    """Local phase of the distributed kNN: rows of k (index, rank) slots, padded with (-1, -1) / +inf."""
    abx, _lib, L, space = ctx
    import oracle
    q, rank = 700, 5
    pts = clouds.uniform01(21, max(n, 1))[:n]
    qs = clouds.uniform01(22, q)
    bvh = abx.BoundingVolumeHierarchy(space, torch.from_numpy(pts).cuda().reshape(-1, 3), abx.POINT)
    vals = torch.full((q * k, 2), -9, dtype=torch.int32, device="cuda")
    dist = torch.full((q * k,), -1.0, dtype=torch.float32, device="cuda")
    missing = C.c_int64(-1)
    d_q = torch.from_numpy(qs).cuda()
    _lib.check(L.abx_dist_nearest_pairs(bvh._h, space.handle, _ptr(d_q), q, k, rank, _ptr(vals), _ptr(dist),
                                        C.byref(missing)))
    row = min(n, k)
    assert missing.value == q * (k - row)
    v = vals.cpu().numpy().reshape(q, k, 2)
    d = dist.cpu().numpy().reshape(q, k)
    assert (v[:, row:] == -1).all() and np.isinf(d[:, row:]).all()
    if n:
        roff, ridx, rd = oracle.Tree(pts).nearest_crs(qs, k)
        assert np.array_equal(d[:, :row], rd.reshape(q, row))
        assert (v[:, :row, 1] == rank).all()
        dd = np.linalg.norm(pts[v[:, :row, 0]].astype(np.float64) - qs[:, None, :], axis=2)
        assert np.allclose(dd, d[:, :row], rtol=1e-5, atol=1e-7)


def test_pair_with_rank(ctx):
    abx, _lib, L, space = ctx
    idx = torch.arange(1000, dtype=torch.int32, device="cuda") * 3
    out = torch.empty((1000, 2), dtype=torch.int32, device="cuda")
    _lib.check(L.abx_dist_pair_with_rank(space.handle, _ptr(idx), 1000, 6, _ptr(out)))
    assert torch.equal(out[:, 0], idx) and bool((out[:, 1] == 6).all())


def test_cross_stream_free_and_result_ownership(ctx):
    """abx_free under another stream than the producing one orders the block's reuse after its last user;
    host-predicate queries return buffers the caller owns (a second query does not overwrite the first)."""
    abx, _lib, L, space = ctx
    pts = torch.from_numpy(clouds.uniform01(31, 50_000)).cuda()
    bvh = abx.BoundingVolumeHierarchy(space, pts)
    other = abx.ExecutionSpace(torch.cuda.Stream())
    off, idx, nnz = C.c_void_p(), C.c_void_p(), C.c_int64()
    sp = torch.cat([pts[:20_000], torch.full((20_000, 1), 0.05, device="cuda")], 1).contiguous()
    ref_idx, ref_off = bvh.query(space, abx.intersects(sp))
    for _ in range(3):
        _lib.check(L.abx_query_spatial_crs(bvh._h, space.handle, 0, _ptr(sp), sp.shape[0], None, _lib.ALLOC_FN(0),
                                           None, C.byref(off), C.byref(idx), C.byref(nnz)))
        _lib.check(L.abx_free(other.handle, idx))  # released under a different stream
        _lib.check(L.abx_free(other.handle, off))
        a, b = bvh.query(other, abx.intersects(sp))  # reuses the blocks on `other`
        assert torch.equal(b, ref_off)
    h1 = bvh.query(space, abx.intersects(sp[:1000].cpu()))
    keep = h1[0].clone()
    h2 = bvh.query(space, abx.intersects(sp[1000:2000].cpu()))
    assert torch.equal(h1[0], keep) and h1[0].data_ptr() != h2[0].data_ptr()
    pool = abx.HostBufferPool()
    p1 = bvh.query(space, abx.intersects(sp[:1000].cpu()), out=pool)
    p2 = bvh.query(space, abx.intersects(sp[:1000].cpu()), out=pool)
    assert p1[0].data_ptr() == p2[0].data_ptr() and torch.equal(p1[0], h1[0])
