"""Rank-count-agnostic DistributedTree cases restated from the reference's MPI tests
(test/tstDistributedTreeSpatial.cpp:32-94,461-548, test/tstDistributedTreeNearest.cpp:64-130,404-452,
examples/distributed_tree/distributed_knn.cpp:62-104).  run_all() is called on every rank of a process
group (gloo + oracle engine on CPU; nccl + CUDA engine on the GPU box)."""
import numpy as np
import torch
import torch.distributed as dist

import arborx_b200.distributed as D
from tests import brute, clouds

F = np.float32


class OracleEngine:
    """Local trees through the CPU oracle (test double for the CUDA engine)."""

    def build(self, values, kind):
        import oracle
        v = values.detach().cpu().numpy().astype(F)
        return oracle.Tree(v.reshape(-1, {0: 3, 1: 6, 2: 9}[kind]), kind)

    def size(self, tree):
        return tree.n

    def bounds(self, tree):
        return torch.from_numpy(tree.bounds())

    def spatial(self, tree, pred_kind, preds):
        p = preds.detach().cpu().numpy().astype(F)
        off, idx = tree.spatial_crs(p, pred_kind)
        return torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(off)

    def nearest(self, tree, pts, k):
        p = pts.detach().cpu().numpy().astype(F)
        off, idx, d = tree.nearest_crs(p, int(k))
        return torch.from_numpy(idx.astype(np.int32)), torch.from_numpy(off), torch.from_numpy(d)


    def merge_rows(self, local_off, local_idx, rank, remote_off, remote_vals):
        lo, li = local_off.cpu().numpy().astype(np.int64), local_idx.cpu().numpy()
        ro, rv = remote_off.cpu().numpy().astype(np.int64), remote_vals.cpu().numpy().reshape(-1, 2)
        q = len(lo) - 1
        off = np.zeros(q + 1, np.int64)
        off[1:] = np.cumsum((lo[1:] - lo[:-1]) + (ro[1:] - ro[:-1]))
        vals = np.zeros((off[-1], 2), np.int32)
        for i in range(q):
            a = off[i]
            nl = lo[i + 1] - lo[i]
            vals[a:a + nl, 0] = li[lo[i]:lo[i + 1]]
            vals[a:a + nl, 1] = rank
            vals[a + nl:off[i + 1]] = rv[ro[i]:ro[i + 1]]
        return torch.from_numpy(vals), torch.from_numpy(off.astype(np.int32))


class _Pred:
    def __init__(self, tag, kind, data, k=None):
        self.tag, self.kind, self.data, self.k = tag, kind, data, k


def rows(vals, off):
    vals = vals.cpu().numpy()
    off = off.cpu().numpy()
    return [sorted(map(tuple, vals[off[i]:off[i + 1]].tolist())) for i in range(len(off) - 1)]


def run_all(make_engine, device, space=None):
    """Protocol model (torch.distributed + injected engine) on every rank of the default process group."""
    comm = dist.group.WORLD
    make_tree = lambda values, kind=None: D.DistributedTree(comm, space, values, kind, engine=make_engine())
    return run_cases(dist.get_rank(), dist.get_world_size(), make_tree, device, space)


def expand_compact(res, rank, with_dist):
    """Host (compact) result of the native tree -> ((index, rank) pairs, offsets[, distances])."""
    idx, off = res[0], res[1]
    rpos, rrank = res[-2], res[-1]
    vals = torch.stack([idx, torch.full_like(idx, rank)], 1)
    vals[rpos.long(), 1] = rrank
    assert bool((rpos[1:] > rpos[:-1]).all())
    return (vals, off) + ((res[2],) if with_dist else ())


def run_cases(rank, size, make_tree, device, space=None, check_host=False):
    """Rank-count-agnostic cases; make_tree(values, kind=None) builds this rank's DistributedTree.
    check_host (native tree): every query is repeated through the host-buffer entry points and the compact
    result must expand to the same rows."""
    class _Tree:
        def __init__(self, values, kind=None):
            self.t = make_tree(values, kind)

        def size(self):
            return self.t.size()

        def empty(self):
            return self.t.empty()

        def query(self, space, pred, return_distances=False):
            out = self.t.query(space, pred, return_distances=return_distances)
            if check_host:
                hp = _Pred(pred.tag, pred.kind, pred.data.cpu(), pred.k)
                hres = expand_compact(self.t.query(space, hp, return_distances=return_distances), rank,
                                      return_distances)
                assert rows(hres[0], hres[1]) == rows(out[0], out[1])
                assert torch.equal(hres[1], out[1].cpu())
                if return_distances and pred.tag == "nearest":
                    assert torch.equal(hres[2], out[2].cpu())
            return out

    n = 4
    T = lambda a: torch.as_tensor(np.asarray(a, F)).to(device)
    n = 4

    # ---- hello world, spatial (tstDistributedTreeSpatial.cpp:32-94) ----
    pts = np.array([[i / n + rank, 0, 0] for i in range(n)], F)
    tree = _Tree(T(pts))
    assert tree.size() == n * size and not tree.empty()
    q = np.array([[0.5 + size - 1 - rank, 0, 0, 0.5]], F)
    vals, off = tree.query(space, _Pred("spatial", D.SPHERE_PRED, T(q)))
    expect = [(n - 1 - i, size - 1 - rank) for i in range(n)]
    if rank > 0:
        expect.append((0, size - rank))
    assert rows(vals, off) == [sorted(expect)], (rank, rows(vals, off), expect)

    # ---- hello world, nearest (tstDistributedTreeNearest.cpp:64-130) ----
    k = 3 if rank < size - 1 else 2
    qn = np.array([[0.0 + size - 1 - rank, 0, 0]], F)
    vals, off, d = tree.query(space, _Pred("nearest", D.POINT_PRED, T(qn), k), return_distances=True)
    if rank < size - 1:
        expect = [(0, size - 1 - rank), (n - 1, size - 2 - rank), (1, size - 1 - rank)]
    else:
        expect = [(0, size - 1 - rank), (1, size - 1 - rank)]
    assert rows(vals, off) == [sorted(expect)], (rank, rows(vals, off), expect)
    assert np.all(np.diff(d.cpu().numpy()) >= 0)

    # ---- non-approximate nearest neighbours (:404-452), box primitives ----
    boxes = np.array([[rank, 0, 0, rank, 0, 0], [rank + 1, 1, 1, rank + 1, 1, 1]], F)
    tb = _Tree(T(boxes), D.BOX)
    assert tb.size() == 2 * size
    qn = np.array([[(size - 1 - rank) + 0.75, 0, 0]], F)
    vals, off = tb.query(space, _Pred("nearest", D.POINT_PRED, T(qn), 1))
    assert rows(vals, off) == [[(0, size - rank - (1 if rank == 0 else 0))]], (rank, rows(vals, off))

    # ---- distributed_knn example (examples/distributed_tree/distributed_knn.cpp:62-104) ----
    pe = np.array([[rank, rank, rank], [rank + .5, rank + .5, rank + .5]], F)
    te = _Tree(T(pe))
    vals, off = te.query(space, _Pred("nearest", D.POINT_PRED, T(pe), 3))
    if rank == 0 and size >= 2:
        r = rows(vals, off)
        assert list(off.cpu().numpy()) == [0, 3, 6]
        assert r[0] == sorted([(0, 0), (1, 0), (0, 1)]) and r[1] == sorted([(1, 0), (0, 0), (0, 1)]), r

    # ---- empty tree and partially empty ranks (tstDistributedTreeSpatial.cpp:96-190) ----
    tz = _Tree(T(np.zeros((0, 3), F)), D.POINT)
    assert tz.empty() and tz.size() == 0
    vals, off = tz.query(space, _Pred("spatial", D.SPHERE_PRED, T([[0, 0, 0, 1], [1, 1, 1, 2]])))
    assert list(off.cpu().numpy()) == [0, 0, 0] and vals.shape[0] == 0
    vals, off = tz.query(space, _Pred("nearest", D.POINT_PRED, T([[0, 0, 0]]), 3))
    assert list(off.cpu().numpy()) == [0, 0]
    only0 = pts if rank == 0 else np.zeros((0, 3), F)
    t0 = _Tree(T(only0), D.POINT)
    assert t0.size() == n
    vals, off = t0.query(space, _Pred("nearest", D.POINT_PRED, T([[0.3, 0, 0]]), 2))
    assert rows(vals, off) == [[(1, 0), (2, 0)]], rows(vals, off)
    vals, off = t0.query(space, _Pred("spatial", D.SPHERE_PRED, T([[0.3, 0, 0, 0.06]])))
    assert rows(vals, off) == [[(1, 0)]]

    # ---- random clouds against a gathered single tree (tstDistributedTreeSpatial.cpp:461-548) ----
    n_loc, q_loc = 3000, 400
    all_pts = [clouds.uniform01(100 + r, n_loc) + F(0.6) * F(r) for r in range(size)]  # overlapping slabs in x
    mine = all_pts[rank]
    tr = _Tree(T(mine))
    qs = (clouds.uniform01(500 + rank, q_loc) * F(0.6 * (size - 1) + 1.0)).astype(F)
    qs[:, 1:] = clouds.uniform01(600 + rank, q_loc)[:, 1:]
    glob = np.concatenate(all_pts)
    owner = np.repeat(np.arange(size), n_loc)
    local_index = np.tile(np.arange(n_loc), size)
    spheres = np.concatenate([qs, np.full((q_loc, 1), 0.07, F)], 1).astype(F)
    vals, off = tr.query(space, _Pred("spatial", D.SPHERE_PRED, T(spheres)))
    mask = brute.spheres_vs_points(spheres, glob)
    expect = [sorted((int(local_index[j]), int(owner[j])) for j in np.nonzero(m)[0]) for m in mask]
    assert rows(vals, off) == expect
    for kk in (1, 7):
        vals, off, d = tr.query(space, _Pred("nearest", D.POINT_PRED, T(qs), kk), return_distances=True)
        Dm = brute.dist_point_point(qs, glob)
        offn = off.cpu().numpy()
        assert list(offn) == [kk * i for i in range(q_loc + 1)]
        dn = d.cpu().numpy().reshape(q_loc, kk)
        assert np.array_equal(dn, np.sort(Dm, 1)[:, :kk])
        v = vals.cpu().numpy().reshape(q_loc, kk, 2)
        gid = v[:, :, 1] * n_loc + v[:, :, 0]
        assert np.array_equal(np.take_along_axis(Dm, gid, 1), dn)
    return True
