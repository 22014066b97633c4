"""Distributed DBSCAN: ArborX::Experimental::dbscan(comm, space, primitives, eps, core_min_size, labels, params)
(cluster/ArborX_DistributedDBSCAN.hpp:29-190, cluster/detail/ArborX_DistributedDBSCANHelpers.hpp).

One process per GPU; the points are sharded as the caller sharded them (spatially compact shards keep the
halo small).  The path has three exchange steps and nothing else crosses ranks:

  1. halo out   all-gather of the rank boxes, then one all-to-all-v of the points that lie within the
                ghost distance of another rank's box (eps for core_min_size == 2, nextafter(2 eps) otherwise:
                a ghost's own core status must be decidable from what its host rank sees),
  2. local      ArborX::dbscan on local + ghost points (libabx.so, abx_dbscan) -- the hot kernel,
  3. labels     local labels -> global ids (rank offset + index), ghost labels back to the owners
                (all-to-all-v), owners derive merge pairs (label -> smaller label) for their core points
                that carry several labels, all-gather-v of the merge pairs, every rank flattens its labels
                through the sorted pair table.

The product path is C++ (abx_dist_dbscan_points3f in libabx.so: kernels + grouped NCCL exchanges, four blocking
points); `dbscan(comm, space, points, ...)` binds it.  With `engine=` the same protocol runs as torch tensor code over
torch.distributed with an injectable local engine: the model the CPU tests exercise (gloo + the oracle)."""
import numpy as np
import torch
import torch.distributed as dist

from .distributed import CudaEngine, SPHERE_PRED, _alltoallv, route_generic

POINT = 0


class CudaDBSCANEngine(CudaEngine):
    """Local DBSCAN and neighbour counts on this rank's GPU through libabx.so."""

    def dbscan(self, pts, eps, core_min_size, params):
        return self.abx.dbscan(self.space, pts, eps, core_min_size, params)

    def count_within(self, pts, query_ids, eps, limit):
        """For the points pts[query_ids]: min(limit, number of points of pts within eps) (CountUpToN)."""
        tree = self.abx.BoundingVolumeHierarchy(self.space, pts, POINT)
        q = pts[query_ids]
        spheres = torch.cat([q, torch.full((q.shape[0], 1), float(np.float32(eps)), device=q.device)], 1).contiguous()
        return tree.count(self.space, self.abx.intersects(spheres), limit=int(limit))


def _ghost_distance(eps, core_min_size):
    """DistributedDBSCAN.hpp:77-84: eps for the connected-components case, nextafter(2 eps) otherwise."""
    e = np.float32(eps)
    if core_min_size == 2:
        return float(e)
    return float(np.nextafter(np.float32(2) * e, np.float32(10) * e, dtype=np.float32))


def _sort_and_filter(pairs):
    """sortAndFilterMergePairs (DistributedDBSCANHelpers.hpp:501-561): per `from` keep the lowest `to`, and
    link every other `to` of that `from` to the lowest one."""
    if pairs.shape[0] == 0:
        return pairs
    p = torch.unique(pairs, dim=0)  # lexicographic (from, to), duplicates dropped
    first = torch.ones(p.shape[0], dtype=torch.bool, device=p.device)
    first[1:] = p[1:, 0] != p[:-1, 0]
    group = torch.cumsum(first.long(), 0) - 1
    lowest = p[first, 1][group]
    extra = torch.stack([p[~first, 1], lowest[~first]], 1)
    out = torch.cat([p[first], extra], 0)
    return torch.unique(out, dim=0)


def _relabel(pairs, labels):
    """relabel (DistributedDBSCANHelpers.hpp:619-660): follow label -> to while the label is a `from`."""
    if pairs.shape[0] == 0:
        return labels
    frm, to = pairs[:, 0].contiguous(), pairs[:, 1].contiguous()
    m = frm.shape[0]
    labels = labels.clone()
    active = torch.arange(labels.shape[0], device=labels.device)
    while active.numel():
        cur = labels[active]
        pos = torch.searchsorted(frm, cur)  # first pair of that `from`: its lowest `to`
        hit = (pos < m) & (frm[pos.clamp(max=m - 1)] == cur)
        active = active[hit]
        labels[active] = to[pos[hit]]
    return labels


def _dbscan_native(comm, space, points, eps, core_min_size, parameters):
    """The C++ path: abx_dist_dbscan_points3f (halo routing kernel, grouped NCCL exchanges, abx::dbscan on local +
    ghost points, merge-pair kernels; arborx_b200/csrc/abx_dist.cu)."""
    import ctypes as C

    from . import DBSCANParameters, _lib
    from .distributed import Communicator
    if not isinstance(comm, Communicator):
        comm = Communicator.from_process_group(comm)
    p = parameters or DBSCANParameters()
    x = points.to(device=space.device, dtype=torch.float32).reshape(-1, 3).contiguous()
    labels = torch.empty(x.shape[0], dtype=torch.int64, device=space.device)
    with torch.cuda.stream(space.stream):
        _lib.check(_lib.lib().abx_dist_dbscan_points3f(comm._h, space.handle, C.c_void_p(x.data_ptr()), x.shape[0],
                                                       float(eps), int(core_min_size), p._implementation, p._algorithm,
                                                       C.c_void_p(labels.data_ptr())))
    return labels


def dbscan(comm, space, points, eps, core_min_size, parameters=None, engine=None):
    """-> labels [n_local] int64: global id (rank offset + local index) of the cluster's representative
    point, -1 for noise; equal labels across ranks mean the same cluster.  Collective.  engine=None: the C++ path of
    libabx.so (`comm`: a torch.distributed group or a Communicator); with an injected engine: the torch.distributed
    protocol model below (CPU tests: gloo + the oracle as the local engine)."""
    if engine is None:
        return _dbscan_native(comm, space, points, eps, core_min_size, parameters)
    if not (eps > 0):
        raise ValueError("eps must be positive")
    if core_min_size < 2:
        raise ValueError("core_min_size must be at least 2")
    rank, R = dist.get_rank(comm), dist.get_world_size(comm)
    pts = points.to(torch.float32).reshape(-1, 3).contiguous()
    dev = pts.device
    n_local = pts.shape[0]
    special = core_min_size == 2

    # ---- step 1: ghosts (forwardNeighbors, Helpers.hpp:285-356) ------------------------------------
    if n_local:
        box = torch.cat([pts.min(0).values, pts.max(0).values])
    else:  # empty box: nothing is near it
        big = torch.finfo(torch.float32).max
        box = torch.tensor([big] * 3 + [-big] * 3, dtype=torch.float32, device=dev)
    boxes = [torch.empty_like(box) for _ in range(R)]
    dist.all_gather(boxes, box, group=comm)
    rank_boxes = torch.stack(boxes)
    sizes = torch.tensor([n_local], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(R)]
    dist.all_gather(all_sizes, sizes, group=comm)
    rank_offsets = torch.zeros(R + 1, dtype=torch.int64, device=dev)
    rank_offsets[1:] = torch.cumsum(torch.cat(all_sizes), 0)

    g_eps = _ghost_distance(eps, core_min_size)
    radius = torch.full((max(n_local, 1),), g_eps, dtype=torch.float32, device=dev)
    if hasattr(engine, "route") and n_local:
        qid_s, send_counts = engine.route(SPHERE_PRED, pts, rank_boxes, rank, radius, 1)
    else:
        qid_s, send_counts = route_generic(SPHERE_PRED, torch.cat([pts, radius[:n_local].unsqueeze(1)], 1),
                                           rank_boxes, rank)
    rows = torch.cat([pts[qid_s].contiguous().view(torch.int32), qid_s.to(torch.int32).unsqueeze(1)], 1)
    got, recv_counts = _alltoallv(comm, rows, send_counts)
    ghost_points = got[:, :3].clone(memory_format=torch.contiguous_format).view(torch.float32)
    ghost_ids = got[:, 3].long()
    ghost_ranks = torch.repeat_interleave(torch.arange(R, device=dev), torch.tensor(recv_counts, device=dev))
    n_ghost = ghost_points.shape[0]

    # ---- step 2: local DBSCAN on local + ghost points ------------------------------------------------
    unified = torch.cat([pts, ghost_points], 0).contiguous()
    if unified.shape[0]:
        local_labels = engine.dbscan(unified, eps, core_min_size, parameters).long().to(dev)
    else:  # a rank without points and without neighbours still takes part in the collectives below
        local_labels = torch.empty(0, dtype=torch.int64, device=dev)

    # ---- step 3: local -> global labels (convertLocalToGlobal, Helpers.hpp:132-175) ------------------
    is_ghost_label = local_labels >= n_local
    gl = (local_labels - n_local).clamp(min=0)
    if n_ghost:
        from_ghost = rank_offsets[ghost_ranks[gl.clamp(max=n_ghost - 1)]] + ghost_ids[gl.clamp(max=n_ghost - 1)]
    else:
        from_ghost = torch.zeros_like(local_labels)
    glob = torch.where(is_ghost_label, from_ghost, rank_offsets[rank] + local_labels)
    glob = torch.where(local_labels < 0, torch.full_like(glob, -1), glob)
    labels = glob[:n_local].clone()
    ghost_labels = glob[n_local:]

    # ---- step 4: ghost labels back to their owners (noise is not sent) -------------------------------
    keep = ghost_labels != -1
    back_counts = [int(c) for c in torch.bincount(ghost_ranks[keep], minlength=R).tolist()] if n_ghost else [0] * R
    gl64 = ghost_labels[keep].contiguous()
    back_rows = torch.cat([gl64.view(torch.int32).view(-1, 2), ghost_ids[keep].to(torch.int32).unsqueeze(1)], 1)
    back, _ = _alltoallv(comm, back_rows, back_counts)
    b_ids = back[:, 2].long()
    b_labels = back[:, :2].contiguous().clone(memory_format=torch.contiguous_format).view(torch.int64).view(-1)
    order = torch.argsort(b_ids, stable=True)
    b_ids, b_labels = b_ids[order], b_labels[order]

    # ---- step 5: merge pairs for multi-labelled points (computeMergePairs, Helpers.hpp:420-499) ------
    pairs = torch.empty((0, 2), dtype=torch.int64, device=dev)
    if b_ids.numel():
        uid, inv, cnt = torch.unique_consecutive(b_ids, return_inverse=True, return_counts=True)
        local = labels[uid]
        valid = local != -1
        if special:
            core = torch.ones_like(valid)  # every point with a neighbour is core
        else:
            core = engine.count_within(unified, uid, eps, core_min_size).to(dev) >= core_min_size
        big = torch.iinfo(torch.int64).max
        gmin = torch.full((uid.shape[0],), big, dtype=torch.int64, device=dev).scatter_reduce(0, inv, b_labels, "amin")
        first_pos = torch.cumsum(cnt, 0) - cnt
        multi = (cnt + valid.long()) >= 2
        # a border point without a local label takes one of the imported ones
        border_fix = multi & ~core & ~valid
        labels[uid[border_fix]] = b_labels[first_pos[border_fix]]
        take = multi & core
        min_label = torch.where(valid, torch.minimum(local, gmin), gmin)
        own = take & valid & (local != min_label)
        own_pairs = torch.stack([local[own], min_label[own]], 1)
        labels[uid[own]] = min_label[own]
        e_take = take[inv] & (b_labels != min_label[inv])
        ghost_pairs = torch.stack([b_labels[e_take], min_label[inv][e_take]], 1)
        pairs = _sort_and_filter(torch.cat([own_pairs, ghost_pairs], 0))

    # ---- step 6: all-gather-v of the merge pairs -----------------------------------------------------
    npairs = torch.tensor([pairs.shape[0]], dtype=torch.int64, device=dev)
    all_np = [torch.empty_like(npairs) for _ in range(R)]
    dist.all_gather(all_np, npairs, group=comm)
    counts = [int(x.item()) for x in all_np]
    mx = max(counts)
    if mx:
        padded = torch.zeros((mx, 2), dtype=torch.int64, device=dev)
        padded[:pairs.shape[0]] = pairs
        gathered = [torch.empty_like(padded) for _ in range(R)]
        dist.all_gather(gathered, padded, group=comm)
        global_pairs = _sort_and_filter(torch.cat([g[:c] for g, c in zip(gathered, counts)], 0))
        # ---- step 7: flatten ---------------------------------------------------------------------
        labels = _relabel(global_pairs, labels)
    return labels
