"""Rank-count-agnostic cases for the distributed DBSCAN (cluster/ArborX_DistributedDBSCAN.hpp; the reference
ships no unit test for it, only benchmarks/cluster/distributed_dbscan.cpp with its verifier).  A global cloud
is split over the ranks (spatial slabs, and a scattered split: any partition must give the same clustering),
every rank runs arborx_b200.distributed_dbscan.dbscan, the labels are gathered and checked against the
single-process oracle: identical core partition and noise set, border points validated by the reference's
verifier (oracle.dbscan_verify).  run_all() is called on every rank (gloo + oracle engine on CPU; nccl +
CUDA engine on the GPU box)."""
import numpy as np
import torch
import torch.distributed as dist

import oracle
from arborx_b200 import DBSCANParameters
from arborx_b200.distributed_dbscan import dbscan as dist_dbscan
from tests import clouds

F = np.float32


class OracleDBSCANEngine:
    """CPU oracle as the local engine (test double for CudaDBSCANEngine)."""

    def dbscan(self, pts, eps, core_min_size, params):
        p = params or DBSCANParameters()
        x = pts.detach().cpu().numpy().astype(F)
        return torch.from_numpy(oracle.dbscan(x, eps, core_min_size, p._implementation, p._algorithm))

    def count_within(self, pts, query_ids, eps, limit):
        x = pts.detach().cpu().numpy().astype(F)
        ids = query_ids.detach().cpu().numpy()
        tree = oracle.Tree(x, 0)
        spheres = np.concatenate([x[ids], np.full((len(ids), 1), F(eps), F)], 1)
        off, _ = tree.spatial_crs(spheres, 0)
        return torch.from_numpy(np.minimum(np.diff(off), limit).astype(np.int32))


def _cloud(seed, n):
    """Blobs of different density + a uniform background, in [0, 10]^3."""
    c = clouds.uniform01(seed, 12) * F(10)
    parts = []
    per = n // 16
    for i in range(12):
        g = (clouds.uniform01(seed + 100 + i, per) - F(0.5)) * F(0.3 + 0.15 * (i % 4))
        parts.append(c[i] + g)
    parts.append(clouds.uniform01(seed + 7, n - 12 * per) * F(10))
    return np.concatenate(parts).astype(F)


def _split(xyz, world, how, seed):
    n = len(xyz)
    if how == "slabs":
        order = np.argsort(xyz[:, 0], kind="stable")
        bounds = [n * r // world for r in range(world + 1)]
        return [order[bounds[r]:bounds[r + 1]] for r in range(world)]
    if how == "scattered":
        owner = (clouds.uniform01(seed + 55, n)[:, 0] * world).astype(np.int64).clip(0, world - 1)
        return [np.nonzero(owner == r)[0] for r in range(world)]
    # "lopsided": rank 0 owns almost nothing, the last rank may be empty
    owner = np.minimum((xyz[:, 1] / F(10) * (world + 1)).astype(np.int64), world - 1)
    owner[:3] = 0
    owner[owner == world - 1] = max(world - 2, 0)
    return [np.nonzero(owner == r)[0] for r in range(world)]


def _canonical(labels, members):
    """label -> smallest index among its `members` (core points: border points may legitimately join any
    adjacent cluster, so they must not name it); noise and labels without such a member stay -1."""
    labels = np.asarray(labels, np.int64)
    out = np.full(len(labels), -1, np.int64)
    ok = (labels >= 0) & members
    if ok.any():
        uniq, inv = np.unique(labels[ok], return_inverse=True)
        first = np.full(len(uniq), len(labels), np.int64)
        np.minimum.at(first, inv, np.nonzero(ok)[0])
        out[ok] = first[inv]
    return out


def run_all(make_engine, device, space=None, n=4000):
    """The torch.distributed protocol model with an injected engine, on every rank of the default group."""
    rank, world = dist.get_rank(), dist.get_world_size()
    engine = make_engine() if space is None else make_engine(space)

    def gather(a):
        out = [None] * world
        dist.all_gather_object(out, a)
        return out

    run = lambda pts, eps, minpts, params: dist_dbscan(dist.group.WORLD, space, pts, eps, minpts, params, engine=engine)
    return run_cases(rank, world, run, gather, device, n)


def run_cases(rank, world, run, gather, device, n=4000):
    """run(points, eps, minpts, params) -> this rank's labels; gather(array) -> the arrays of all ranks."""
    ran = 0
    for seed, eps, minpts, impl, how in [(1, 0.12, 2, 0, "slabs"), (1, 0.12, 5, 0, "slabs"), (2, 0.2, 3, 1, "slabs"),
                                         (3, 0.15, 2, 1, "scattered"), (3, 0.15, 4, 0, "scattered"),
                                         (4, 0.3, 5, 1, "lopsided"), (5, 0.05, 2, 0, "lopsided"),
                                         (6, 0.6, 10, 0, "slabs")]:
        xyz = _cloud(seed, n)
        parts = _split(xyz, world, how, seed)
        mine = parts[rank]
        pts = torch.from_numpy(xyz[mine]).to(device)
        labels = run(pts, eps, minpts, DBSCANParameters(impl, 0))
        assert labels.shape == (len(mine),) and labels.dtype == torch.int64
        gathered = gather(labels.cpu().numpy())
        # global id = rank offset + local index = position in the rank-ordered concatenation
        order = np.concatenate(parts)
        xyz_cat = xyz[order]
        lab_cat = np.concatenate(gathered)
        assert lab_cat.max(initial=-1) < len(order)
        ref, core = oracle.dbscan(xyz_cat, eps, minpts, impl=0, algo=0, return_core=True)
        core = core.astype(bool)
        got = _canonical(lab_cat, core)
        want = _canonical(ref, core)
        assert np.array_equal(got[core], want[core]), "core partition differs (%s, eps=%g, minpts=%d)" % (how, eps, minpts)
        assert np.array_equal(lab_cat[~core] == -1, ref[~core] == -1), "noise set differs"
        # border points: the reference's verifier on compact int32 labels
        _, compact = np.unique(lab_cat, return_inverse=True)
        compact = np.where(lab_cat >= 0, compact, -1).astype(np.int32)
        assert oracle.dbscan_verify(xyz_cat, eps, minpts, compact, 0) == 0
        ran += 1
    return ran
