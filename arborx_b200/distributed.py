"""DistributedTree -- Python binding of ArborX::DistributedTree for one process (or host thread) per GPU.

The product path is C++ (arborx_b200/csrc/abx_dist.cu behind include/abx.h: abx_comm_*, abx_dist_create,
abx_dist_query_spatial_crs, abx_dist_query_nearest_crs and their *_host forms): bottom tree, all-gather of the
rank boxes, routing kernel, two grouped NCCL exchanges and the merge kernels, two blocking points per query.
`DistributedTree(comm, space, values)` binds it; `comm` is a torch.distributed process group (a NCCL
communicator for the library is bootstrapped over it once and cached) or a `Communicator`.

Reference: distributed/ArborX_DistributedTree.hpp:33-252, detail/ArborX_DistributedTreeSpatial.hpp:31-60,
detail/ArborX_DistributedTreeNearest.hpp:41-218, detail/ArborX_DistributedTreeUtils.hpp:52-342,
detail/ArborX_Distributor.hpp:276-440.

`ProtocolModelTree` is the reference-shaped exchange (every query forwarded through the top tree, two-phase
kNN) written with torch tensor ops over torch.distributed; its local engine is injectable, so the CPU tests run it
with gloo and the oracle as the engine (world sizes 1-3).  It is a model of the protocol for those tests, not
the product path; `DistributedTree(..., engine=...)` selects it.

Values returned by queries are (index, rank) pairs (int32 [nnz, 2]): the reference returns user values or
`{index, rank}` from a callback (examples/distributed_tree/distributed_knn.cpp:62-104).
"""
import ctypes as C

import torch
import torch.distributed as dist

POINT, BOX, TRIANGLE = 0, 1, 2
SPHERE_PRED, BOX_PRED, POINT_PRED = 0, 1, 2
_PRIM_STRIDE = {POINT: 3, BOX: 6, TRIANGLE: 9}


# ------------------------------------------------------------------------- product path ----
class Communicator:
    """abx_comm: the communicator the C++ DistributedTree exchanges over (NCCL, or an in-process group of
    host threads sharing one GPU for single-GPU tests)."""
    _by_group = {}

    def __init__(self, handle):
        self._h = handle

    @property
    def rank(self):
        from . import _lib
        return _lib.lib().abx_comm_rank(self._h)

    @property
    def size(self):
        from . import _lib
        return _lib.lib().abx_comm_size(self._h)

    @classmethod
    def from_process_group(cls, group=None):
        """NCCL communicator over the ranks of a torch.distributed group (collective on first use, cached):
        rank 0 creates the unique id, the group broadcasts it, every rank joins."""
        from . import _lib
        key = id(group) if group is not None else 0
        comm = cls._by_group.get(key)
        if comm is not None:
            return comm
        L = _lib.lib()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [None]
        if rank == 0:
            buf = C.create_string_buffer(128)
            _lib.check(L.abx_comm_unique_id(buf))
            box[0] = buf.raw
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        h = C.c_void_p()
        _lib.check(L.abx_comm_init_rank(box[0], world, rank, C.byref(h)))
        comm = cls(h)
        cls._by_group[key] = comm
        return comm

    @classmethod
    def local_group(cls, n):
        """n communicators of one in-process group (rank r must be driven by its own host thread)."""
        from . import _lib
        arr = (C.c_void_p * n)()
        _lib.check(_lib.lib().abx_comm_create_local(n, arr))
        return [cls(C.c_void_p(arr[r])) for r in range(n)]


class DistributedTree:
    """ArborX::DistributedTree(comm, space, values) (distributed/ArborX_DistributedTree.hpp:129-151).
    Construction and queries are collective over `comm`."""

    def __new__(cls, comm, space, values, kind=None, engine=None):
        if engine is not None:  # protocol model with an injected local engine (CPU tests)
            return ProtocolModelTree(comm, space, values, kind, engine)
        return super().__new__(cls)

    def __init__(self, comm, space, values, kind=None, engine=None):
        from . import _lib
        if not isinstance(comm, Communicator):
            comm = Communicator.from_process_group(comm)
        self.comm, self.space = comm, space
        self.rank, self.world = comm.rank, comm.size
        if not isinstance(values, torch.Tensor):
            values = torch.as_tensor(values, dtype=torch.float32)
        if kind is None:
            kind = {3: POINT, 6: BOX, 9: TRIANGLE}[values.shape[-1]]
        self.kind = kind
        v = values.to(torch.float32).reshape(-1, _PRIM_STRIDE[kind]).contiguous()
        h = C.c_void_p()
        L = _lib.lib()
        with torch.cuda.stream(space.stream):
            fn = L.abx_dist_create if v.is_cuda else L.abx_dist_create_host
            _lib.check(fn(comm._h, space.handle, kind, C.c_void_p(v.data_ptr()), v.shape[0], C.byref(h)))
        self._values = v
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                from . import _lib
                _lib.lib().abx_dist_destroy(h)
            except Exception:
                pass
            self._h = None

    def size(self):
        from . import _lib
        return _lib.lib().abx_dist_size(self._h)

    def empty(self):
        from . import _lib
        return bool(_lib.lib().abx_dist_empty(self._h))

    def bounds(self):
        from . import _lib
        out = (C.c_float * 6)()
        _lib.check(_lib.lib().abx_dist_bounds(self._h, out))
        return torch.tensor(list(out), dtype=torch.float32)

    def query(self, space, predicates, return_distances=False, out=None):
        """Collective.  Device predicates -> (values int32 [nnz, 2] = (index, rank), offsets int32 [q + 1]
        [, distances]) on the device.  Host (CPU tensor) predicates run the host-buffer entry points and
        return the compact form in pinned host memory: (indices int32 [nnz], offsets[, distances], remote_pos,
        remote_rank) -- every index belongs to this rank except indices[remote_pos[j]], owned by
        remote_rank[j].  `out`: HostBufferPool to reuse for the host results."""
        from . import _Allocator, _lib
        L = _lib.lib()
        d = predicates.data
        q = d.shape[0]
        host = not d.is_cuda
        alloc = _Allocator(space.device, pinned_host=host, pool=out if host else None)
        off, vals, dist_p, rpos, rrank = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nnz, nrem = C.c_int64(), C.c_int64()
        nearest = predicates.tag == "nearest"
        want_d = C.byref(dist_p) if (return_distances and nearest) else None
        with torch.cuda.stream(space.stream):
            if host and nearest:
                _lib.check(L.abx_dist_query_nearest_crs_host(self._h, space.handle, C.c_void_p(d.data_ptr()), q,
                                                             int(predicates.k), alloc.fn, None, C.byref(off),
                                                             C.byref(vals), want_d, C.byref(nnz), C.byref(rpos),
                                                             C.byref(rrank), C.byref(nrem)))
            elif host:
                _lib.check(L.abx_dist_query_spatial_crs_host(self._h, space.handle, predicates.kind,
                                                             C.c_void_p(d.data_ptr()), q, alloc.fn, None, C.byref(off),
                                                             C.byref(vals), C.byref(nnz), C.byref(rpos), C.byref(rrank),
                                                             C.byref(nrem)))
            elif nearest:
                _lib.check(L.abx_dist_query_nearest_crs(self._h, space.handle, C.c_void_p(d.data_ptr()), q,
                                                        int(predicates.k), alloc.fn, None, C.byref(off), C.byref(vals),
                                                        want_d, C.byref(nnz)))
            else:
                _lib.check(L.abx_dist_query_spatial_crs(self._h, space.handle, predicates.kind,
                                                        C.c_void_p(d.data_ptr()), q, alloc.fn, None, C.byref(off),
                                                        C.byref(vals), C.byref(nnz)))
        res = alloc.out
        alloc.out, alloc.fn = {}, None
        dev = "cpu" if host else space.device
        empty_i = lambda: torch.empty(0, dtype=torch.int32, device=dev)
        offsets = res[0]
        values = res.get(1, empty_i())
        if not host:
            values = values.view(-1, 2)
        outp = (values, offsets)
        if return_distances:
            outp += (res.get(2, torch.empty(0, dtype=torch.float32, device=dev)),)
        if host:
            outp += (res.get(3, empty_i()), res.get(4, empty_i()))
        return outp


# ------------------------------------------------------- local engine + tensor helpers ----
class CudaEngine:
    """Local trees on this rank's GPU through libabx.so (used by the distributed DBSCAN driver and, with the
    protocol model, by scripts that cross-check it against the C++ path)."""

    def __init__(self, space):
        import arborx_b200 as abx
        self.abx = abx
        self.space = space

    def build(self, values, kind):
        return self.abx.BoundingVolumeHierarchy(self.space, values, kind)

    def size(self, tree):
        return tree.size()

    def bounds(self, tree):
        return tree.bounds()

    def spatial(self, tree, pred_kind, preds):
        idx, off = tree.query(self.space, self.abx.intersects(preds, pred_kind))
        return idx, off

    def nearest(self, tree, pts, k):
        idx, off, d = tree.query(self.space, self.abx.nearest(pts, int(k)), return_distances=True)
        return idx, off, d

    def route(self, kind, data, rank_boxes, rank, radius=None, radius_stride=1):
        """-> (query ids grouped by destination rank [F] int64, send_counts list[R]); self is never a destination.
        radius (optional, spheres): data holds points and predicate i has radius radius.view(-1)[i * radius_stride]."""
        from . import _lib
        L = _lib.lib()
        dev = data.device
        R = rank_boxes.shape[0]
        q = data.shape[0]
        with torch.cuda.stream(self.space.stream):
            boxes = rank_boxes.to(device=dev, dtype=torch.float32).contiguous()
            counts = torch.empty(R, dtype=torch.int32, device=dev)
            d = data.contiguous()
            rp = C.c_void_p(radius.data_ptr()) if radius is not None else None
            _lib.check(L.abx_dist_route_count(self.space.handle, kind, C.c_void_p(d.data_ptr()), q, rp, radius_stride,
                                              C.c_void_p(boxes.data_ptr()), R, int(rank), C.c_void_p(counts.data_ptr())))
            send_counts = counts.tolist()
            total = sum(send_counts)
            qids = torch.empty(total, dtype=torch.int32, device=dev)
            if total:
                base = torch.tensor([sum(send_counts[:i]) for i in range(R)], dtype=torch.int32, device=dev)
                cursors = torch.empty(R, dtype=torch.int32, device=dev)
                _lib.check(L.abx_dist_route_fill(self.space.handle, kind, C.c_void_p(d.data_ptr()), q, rp, radius_stride,
                                                 C.c_void_p(boxes.data_ptr()), R, int(rank), C.c_void_p(base.data_ptr()),
                                                 C.c_void_p(cursors.data_ptr()), C.c_void_p(qids.data_ptr())))
            return qids.long(), send_counts


def route_generic(kind, data, rank_boxes, rank):
    """Device-agnostic tensor version of CudaEngine.route (used by the CPU protocol tests)."""
    dev = data.device
    boxes = rank_boxes.to(dev)
    R = boxes.shape[0]
    parts = []
    if kind == BOX_PRED:
        qlo, qhi = data[:, 0:3], data[:, 3:6]
    else:
        c = data[:, 0:3]
        if kind == SPHERE_PRED:
            r = data[:, 3]
            r2 = r * r * (1.0 + 1e-4) + 1e-30  # conservative against rounding
    for rk in range(R):
        lo, hi = boxes[rk, 0:3], boxes[rk, 3:6]
        if rk == rank or bool((lo > hi).any()):
            parts.append(torch.empty(0, dtype=torch.int64, device=dev))
            continue
        if kind == BOX_PRED:
            hit = ~(((qlo > hi) | (qhi < lo)).any(1))
        else:
            d = torch.minimum(torch.maximum(c, lo), hi) - c
            d2 = (d * d).sum(1)
            hit = ((d2 <= r2) | torch.isinf(r2)) if kind == SPHERE_PRED else (d2 == 0)
        parts.append(torch.nonzero(hit).flatten())
    return torch.cat(parts), [int(p.shape[0]) for p in parts]


def _alltoallv(comm, rows, send_counts):
    """rows [F, w] (32-bit words) ordered by destination rank, send_counts [R] (host list).
    Returns (recv_rows [G, w], recv_counts list)."""
    R = dist.get_world_size(comm)
    dev = rows.device
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rc = torch.empty(R, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=comm)
    recv_counts = rc.tolist()
    w = rows.shape[1]
    out = torch.empty((sum(recv_counts), w), dtype=rows.dtype, device=dev)
    dist.all_to_all_single(out.view(-1), rows.contiguous().view(-1), [c * w for c in recv_counts],
                           [c * w for c in send_counts], group=comm)
    return out, recv_counts


# ------------------------------------------------------------------- protocol model ----
class ProtocolModelTree:
    """The reference-shaped exchange over torch.distributed with an injectable local engine: top-tree query ->
    bucket by destination -> all-to-all-v of packed records -> bottom-tree query -> all-to-all-v back ->
    segmented sort by query id; kNN in the reference's two phases.  Exercised on CPU (gloo + oracle engine)."""

    def __init__(self, comm, space, values, kind=None, engine=None):
        self.comm = comm
        self.rank = dist.get_rank(comm)
        self.world = dist.get_world_size(comm)
        self.space = space
        self.engine = engine if engine is not None else CudaEngine(space)
        if not isinstance(values, torch.Tensor):
            values = torch.as_tensor(values, dtype=torch.float32)
        if kind is None:
            kind = {3: POINT, 6: BOX, 9: TRIANGLE}[values.shape[-1]]
        self.kind = kind
        self.device = values.device
        # bottom tree (ArborX_DistributedTree.hpp:183-186)
        self._bottom = self.engine.build(values, kind)
        n_local = int(self.engine.size(self._bottom))
        # all-gather rank boxes (:208-227) and sizes (:243-245; 64-bit, separate from the float boxes)
        b = self.engine.bounds(self._bottom).to(torch.float32).to(self.device)
        gathered = [torch.empty_like(b) for _ in range(self.world)]
        dist.all_gather(gathered, b, group=comm)
        self._rank_boxes = torch.stack(gathered).cpu().contiguous()
        nl = torch.tensor([n_local], dtype=torch.int64, device=self.device)
        sizes = [torch.empty_like(nl) for _ in range(self.world)]
        dist.all_gather(sizes, nl, group=comm)
        self._sizes = torch.cat(sizes).cpu()
        self._size = int(self._sizes.sum())
        # replicated top tree over the rank boxes (:227); leaf value = rank
        self._top_tree = None

    @property
    def _top(self):
        if self._top_tree is None:
            self._top_tree = self.engine.build(self._rank_boxes.to(self.device), BOX)
        return self._top_tree

    # ---- ArborX_DistributedTree.hpp:112-127 ----------------------------------------------
    def size(self):
        return self._size

    def empty(self):
        return self._size == 0

    def bounds(self):
        # union of the rank boxes = bounds of the top tree (ArborX_DistributedTree.hpp:122-127)
        b = self._rank_boxes
        valid = (b[:, 0:3] <= b[:, 3:6]).all(1)
        if not bool(valid.any()):
            return self.engine.bounds(self._top)
        return torch.cat([b[valid, 0:3].min(0).values, b[valid, 3:6].max(0).values])

    # ---- query ------------------------------------------------------------------------
    def query(self, space, predicates, return_distances=False):
        """Collective.  -> (values int32 [nnz, 2] = (index, rank), offsets int32 [q + 1][, distances])."""
        data = predicates.data
        q = data.shape[0]
        dev = data.device
        if self.empty():
            # DistributedTreeSpatial.hpp:44-50
            out = (torch.empty((0, 2), dtype=torch.int32, device=dev), torch.zeros(q + 1, dtype=torch.int32, device=dev))
            return out + ((torch.empty(0, dtype=torch.float32, device=dev),) if return_distances else ())
        if predicates.tag == "spatial":
            ranks, off = self.engine.spatial(self._top, predicates.kind, data)
            vals, offsets, _ = self._forward_and_collect(data, ranks.long(), off.long(), ("spatial", predicates.kind))
            return (vals, offsets) + ((torch.empty(0, dtype=torch.float32, device=dev),) if return_distances else ())
        vals, offsets, d = self._nearest(data, int(predicates.k))
        return (vals, offsets, d) if return_distances else (vals, offsets)

    # forwardQueries + bottom query + communicateResultsBack + sort by query id
    # (DistributedTreeUtils.hpp:229-263).  ranks/off: CRS of destination ranks per local query.
    def _forward_and_collect(self, data, ranks, off, what):
        dev = data.device
        q = data.shape[0]
        R = self.world
        counts = off[1:] - off[:-1]
        qid = torch.repeat_interleave(torch.arange(q, device=dev), counts)
        # bucket the export list by destination rank (Distributor::createFromSends, Distributor.hpp:132-197)
        order = torch.argsort(ranks, stable=True)
        send_counts = torch.bincount(ranks, minlength=R).tolist()
        qid_s = qid[order]
        rows = torch.cat([data[qid_s].contiguous().view(torch.int32), qid_s.to(torch.int32).unsqueeze(1)], 1)
        fwd, recv_counts = _alltoallv(self.comm, rows, send_counts)
        stride = data.shape[1]
        fwd_preds = fwd[:, :stride].contiguous().view(torch.float32)
        fwd_ids = fwd[:, stride]
        G = fwd.shape[0]
        # bottom-tree query on the forwarded predicates
        if what[0] == "spatial":
            idx, loff = self.engine.spatial(self._bottom, what[1], fwd_preds)
            dist_bits = None
        else:
            idx, loff, d = self.engine.nearest(self._bottom, fwd_preds, what[1])
            dist_bits = d.contiguous().view(torch.int32)
        idx = idx.to(torch.int32)
        loff = loff.long()
        lcounts = loff[1:] - loff[:-1]
        # results are in CRS order of the forwarded queries, which arrived grouped by source rank:
        # already bucketed by destination
        res_ids = torch.repeat_interleave(fwd_ids, lcounts)
        cols = [idx.unsqueeze(1), res_ids.unsqueeze(1)]
        if dist_bits is not None:
            cols.append(dist_bits.unsqueeze(1))
        back_rows = torch.cat(cols, 1) if G else torch.empty((0, len(cols)), dtype=torch.int32, device=dev)
        seg = torch.cumsum(torch.tensor([0] + recv_counts, device=dev), 0)
        back_counts = (loff[seg[1:]] - loff[seg[:-1]]).tolist()
        got, got_counts = _alltoallv(self.comm, back_rows, back_counts)
        src_rank = torch.repeat_interleave(torch.arange(R, device=dev, dtype=torch.int32),
                                           torch.tensor(got_counts, device=dev), output_size=int(got.shape[0]))
        ids = got[:, 1].long()
        order2 = torch.argsort(ids, stable=True)
        vals = torch.stack([got[:, 0][order2], src_rank[order2]], 1)
        offsets = torch.zeros(q + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(torch.bincount(ids, minlength=q), 0)
        dists = got[:, 2][order2].contiguous().view(torch.float32) if dist_bits is not None else None
        return vals, offsets.to(torch.int32), dists

    # DistributedTreeNearest.hpp:41-218
    def _nearest(self, pts, k):
        dev = pts.device
        q = pts.shape[0]
        if k < 1:
            return (torch.empty((0, 2), dtype=torch.int32, device=dev), torch.zeros(q + 1, dtype=torch.int32, device=dev),
                    torch.empty(0, dtype=torch.float32, device=dev))
        # phase I: nearest rank boxes, truncated once their cumulated sizes reach k (:64-101)
        ranks, off, _ = self.engine.nearest(self._top, pts, k)
        ranks, off = ranks.long(), off.long()
        sizes = self._sizes.to(dev)
        counts = off[1:] - off[:-1]
        row = torch.repeat_interleave(torch.arange(q, device=dev), counts)
        pos = torch.arange(ranks.shape[0], device=dev) - off[row]
        sz = sizes[ranks]
        csum = torch.cumsum(sz, 0)
        before = csum - sz - (csum - sz)[off[row]]  # leaves cumulated before this entry, within the row
        # stop at the first empty tree or once k leaves are cumulated
        empty_before = torch.cumsum((sz == 0).long(), 0)
        empty_before = empty_before - empty_before[off[row]] + (sz[off[row]] == 0).long()
        keep = (before < k) & (empty_before == 0)
        ranks1 = ranks[keep]
        off1 = torch.zeros(q + 1, dtype=torch.int64, device=dev)
        off1[1:] = torch.cumsum(torch.bincount(row[keep], minlength=q), 0)
        _, offd, d1 = self._forward_and_collect(pts, ranks1, off1, ("nearest", k))
        offd = offd.long()
        # k-th smallest distance per query (:114-125)
        nd = offd[1:] - offd[:-1]
        rowd = torch.repeat_interleave(torch.arange(q, device=dev), nd)
        # segmented sort: by (row, distance)
        o = torch.argsort(d1, stable=True)
        o = o[torch.argsort(rowd[o], stable=True)]
        d_sorted = d1[o]
        kth = torch.clamp(torch.minimum(torch.full_like(nd, k), nd) - 1, min=0)
        farthest = torch.where(nd > 0, d_sorted[torch.clamp(offd[:-1] + kth, max=max(d_sorted.shape[0] - 1, 0))]
                               if d_sorted.shape[0] else torch.zeros(q, device=dev), torch.zeros(q, device=dev))
        # phase II: every rank whose box is within that distance (:131-176)
        spheres = torch.cat([pts, farthest.unsqueeze(1).to(torch.float32)], 1).contiguous()
        ranks2, off2 = self.engine.spatial(self._top, SPHERE_PRED, spheres)
        vals, offv, d2 = self._forward_and_collect(pts, ranks2.long(), off2.long(), ("nearest", k))
        offv = offv.long()
        # filterResults (DistributedTreeUtils.hpp:267-342): keep the k smallest per query, ascending
        nv = offv[1:] - offv[:-1]
        rowv = torch.repeat_interleave(torch.arange(q, device=dev), nv)
        o = torch.argsort(d2, stable=True)
        o = o[torch.argsort(rowv[o], stable=True)]
        rank_in_row = torch.arange(o.shape[0], device=dev) - offv[rowv[o]]
        sel = o[rank_in_row < k]
        out_counts = torch.minimum(nv, torch.full_like(nv, k))
        offsets = torch.zeros(q + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(out_counts, 0)
        return vals[sel], offsets.to(torch.int32), d2[sel]
