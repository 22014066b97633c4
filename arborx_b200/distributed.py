"""DistributedTree -- host-side mirror of ArborX::DistributedTree for one process per GPU.

Reference: distributed/ArborX_DistributedTree.hpp:33-252 (ctor: bottom tree, all-gather of rank
boxes, replicated top tree, all-gather of sizes), detail/ArborX_DistributedTreeSpatial.hpp:31-60,
detail/ArborX_DistributedTreeNearest.hpp:41-218 (two-phase kNN), detail/ArborX_DistributedTreeUtils.hpp
(forwardQueries :52-115, communicateResultsBack :153-224, countResults/sort :229-263, filterResults
:267-342).  The reference exchanges with MPI point-to-point, three messages each way
(detail/ArborX_Distributor.hpp:276-440); here every exchange is ONE all-to-all-v of packed 32-bit
records over torch.distributed (NCCL over NVLink on the GPUs; gloo in the CPU protocol tests), preceded
by an all-to-all of the R counts.

The tree work (top-tree queries, bottom-tree queries, kNN) runs in the hand-written CUDA kernels behind
the C ABI; the packing between exchanges is a handful of torch tensor ops (bucket by destination, gather,
segmented sort by query id) -- plumbing around the hot path.  The local engine is injectable so that the
exchange protocol can be exercised on CPU with gloo (tests/ plug the oracle in; the product default is the
CUDA engine and fails without a GPU).

Values returned by queries are (index, rank) pairs (int32 [nnz, 2]): the reference returns user values or
`{index, rank}` from a callback (examples/distributed_tree/distributed_knn.cpp:62-104).
"""
import os
import time

import torch
import torch.distributed as dist

_DEBUG = bool(os.environ.get("ABX_DIST_DEBUG"))
_marks = []


def _mark(name):
    """ABX_DIST_DEBUG=1: host-synchronised section timer (diagnostics only)."""
    if _DEBUG:
        torch.cuda.synchronize()
        _marks.append((name, time.perf_counter()))


def _report(tag):
    if _DEBUG and _marks and dist.get_rank() == 0:
        t0 = _marks[0][1]
        print(tag + ": " + "  ".join("%s=%.2f" % (n, (t - t0) * 1e3) for n, t in _marks[1:]), flush=True)
    _marks.clear()


POINT, BOX, TRIANGLE = 0, 1, 2
SPHERE_PRED, BOX_PRED, POINT_PRED = 0, 1, 2


class CudaEngine:
    """Local trees on this rank's GPU through libabx.so."""

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
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        dev = data.device
        R = rank_boxes.shape[0]
        q = data.shape[0]
        boxes = rank_boxes.to(device=dev, dtype=torch.float32).contiguous()
        counts = torch.empty(R, dtype=torch.int32, device=dev)
        d = data.contiguous()
        with torch.cuda.stream(self.space.stream):
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

    def pair_with_rank(self, idx, rank):
        import ctypes as C
        from . import _lib
        i32 = idx.to(torch.int32).contiguous()
        out = torch.empty((i32.shape[0], 2), dtype=torch.int32, device=i32.device)
        with torch.cuda.stream(self.space.stream):
            _lib.check(_lib.lib().abx_dist_pair_with_rank(self.space.handle, C.c_void_p(i32.data_ptr()), i32.shape[0],
                                                          int(rank), C.c_void_p(out.data_ptr())))
        return out

    def nearest_pairs(self, tree, pts, k, rank):
        """k nearest of every point as (index, rank) pairs [q * row, 2] + distances [q * row], row = min(k, size);
        None when some row could not be filled (the caller takes the general path)."""
        import ctypes as C
        from . import _lib
        q = pts.shape[0]
        row = max(0, min(int(k), tree.size()))
        dev = pts.device
        vals = torch.empty((q * row, 2), dtype=torch.int32, device=dev)
        d = torch.empty(q * row, dtype=torch.float32, device=dev)
        missing = C.c_int64(0)
        p = pts.contiguous()
        with torch.cuda.stream(self.space.stream):
            _lib.check(_lib.lib().abx_dist_nearest_pairs(tree._h, self.space.handle, C.c_void_p(p.data_ptr()), q, int(k),
                                                         int(rank), C.c_void_p(vals.data_ptr()),
                                                         C.c_void_p(d.data_ptr()), C.byref(missing)))
        if missing.value:
            return None
        return vals, d

    def knn_merge(self, ids, cand_vals, cand_d, k, vals, dists):
        """Merge remote candidates (ids ascending) into the k-entry rows of their queries, in place."""
        import ctypes as C
        from . import _lib
        i64 = ids.to(torch.int64).contiguous()
        cv = cand_vals.to(torch.int32).contiguous()
        cd = cand_d.to(torch.float32).contiguous()
        with torch.cuda.stream(self.space.stream):
            _lib.check(_lib.lib().abx_dist_knn_merge(self.space.handle, i64.shape[0], C.c_void_p(i64.data_ptr()),
                                                     C.c_void_p(cv.data_ptr()), C.c_void_p(cd.data_ptr()), int(k),
                                                     C.c_void_p(vals.data_ptr()), C.c_void_p(dists.data_ptr())))

    def merge_sorted(self, local_off, local_idx, rank, remote_ids, remote_vals):
        """Local CRS (index only) + remote records (query id ascending, (index, rank)) -> merged
        (values [nnz, 2], offsets [q + 1]) (abx_dist_merge_sorted)."""
        import ctypes as C
        from . import _lib
        q = local_off.shape[0] - 1
        dev = local_off.device
        m = int(remote_ids.shape[0])
        nnz = int(local_idx.shape[0]) + m
        out_off = torch.empty(q + 1, dtype=torch.int32, device=dev)
        out_vals = torch.empty((nnz, 2), dtype=torch.int32, device=dev)
        lo = local_off.to(torch.int32).contiguous()
        li = local_idx.to(torch.int32).contiguous()
        ri = remote_ids.to(torch.int64).contiguous()
        rv = remote_vals.to(torch.int32).contiguous()
        with torch.cuda.stream(self.space.stream):
            _lib.check(_lib.lib().abx_dist_merge_sorted(self.space.handle, q, C.c_void_p(lo.data_ptr()),
                                                        C.c_void_p(li.data_ptr()), int(rank), m,
                                                        C.c_void_p(ri.data_ptr()), C.c_void_p(rv.data_ptr()),
                                                        C.c_void_p(out_off.data_ptr()), C.c_void_p(out_vals.data_ptr())))
        return out_vals, out_off

    def merge_rows(self, local_off, local_idx, rank, remote_off, remote_vals):
        """CRS rows of local results (index only) + CRS rows of remote results ((index, rank) pairs)
        -> merged (values [nnz, 2], offsets [q + 1]); one kernel pass (abx_dist_merge_crs)."""
        import ctypes as C
        from . import _lib
        q = local_off.shape[0] - 1
        dev = local_off.device
        nnz = int(local_idx.shape[0] + remote_vals.shape[0])
        out_off = torch.empty(q + 1, dtype=torch.int32, device=dev)
        out_vals = torch.empty((nnz, 2), dtype=torch.int32, device=dev)
        lo = local_off.to(torch.int32).contiguous()
        li = local_idx.to(torch.int32).contiguous()
        ro = remote_off.to(torch.int32).contiguous()
        rv = remote_vals.to(torch.int32).contiguous()
        with torch.cuda.stream(self.space.stream):
            _lib.check(_lib.lib().abx_dist_merge_crs(self.space.handle, q, C.c_void_p(lo.data_ptr()),
                                                     C.c_void_p(li.data_ptr()), int(rank), C.c_void_p(ro.data_ptr()),
                                                     C.c_void_p(rv.data_ptr()), C.c_void_p(out_off.data_ptr()),
                                                     C.c_void_p(out_vals.data_ptr())))
        return out_vals, out_off


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


class DistributedTree:
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
        self.force_generic = False  # tests flip this to exercise the reference-shaped exchange
        # ABX_DIST_OVERLAP=1 runs the exchange on a side stream while a helper thread drives the local
        # query.  Measured on 2xB200 (10M/rank): radius phase 10.9 ms without, 68 ms with (host-side
        # contention between the two Python threads) -- off by default.
        self._overlap = os.environ.get("ABX_DIST_OVERLAP", "0") == "1"
        self._side = None
        # bottom tree (ArborX_DistributedTree.hpp:183-186)
        self._bottom = self.engine.build(values, kind)
        n_local = int(self.engine.size(self._bottom))
        # all-gather rank boxes (:208-227) and sizes (:243-245)
        b = self.engine.bounds(self._bottom).to(torch.float32).cpu()
        meta = torch.cat([b, torch.tensor([float(n_local)])]).to(self.device)
        gathered = [torch.empty_like(meta) for _ in range(self.world)]
        dist.all_gather(gathered, meta, group=comm)
        g = torch.stack(gathered).cpu()
        self._rank_boxes = g[:, :6].contiguous()
        self._sizes = g[:, 6].to(torch.int64)
        self._size = int(self._sizes.sum())
        # replicated top tree over the rank boxes (:227); leaf value = rank.  Built on first use: the
        # fast paths route against the R boxes directly and never need it.
        self._top_tree = None

    def _side_space(self):
        if self._side is None:
            import arborx_b200 as abx
            self._side = abx.ExecutionSpace(torch.cuda.Stream(device=self.device, priority=-1))
            self._side_engine = CudaEngine(self._side)
        return self._side

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
        fast = self.world <= 62 and not self.force_generic
        if predicates.tag == "spatial":
            if fast:
                vals, offsets = self._spatial_fast(predicates.kind, data)
            else:
                ranks, off = self.engine.spatial(self._top, predicates.kind, data)
                vals, offsets, _ = self._forward_and_collect(data, ranks.long(), off.long(),
                                                             ("spatial", predicates.kind))
            return (vals, offsets) + ((torch.empty(0, dtype=torch.float32, device=dev),) if return_distances else ())
        k = int(predicates.k)
        if fast and k >= 1 and int(self._sizes.min()) >= k:
            vals, offsets, d = self._nearest_fast(data, k)
        else:
            vals, offsets, d = self._nearest(data, k)
        return (vals, offsets, d) if return_distances else (vals, offsets)

    # ---- fast paths --------------------------------------------------------------------
    # The reference forwards every query through the exchange, including the (vast majority
    # of) queries that only concern the rank they live on.  Here the local tree is queried
    # directly for all local queries, only the queries that also touch OTHER ranks' boxes are
    # packed, exchanged and merged back, so the full-size arrays are touched by the tree
    # kernels and one merge pass only.  Routing may be conservative (a rank that gets a query
    # it has nothing for returns nothing), so it is a plain tensor test against the R boxes.
    def _route(self, kind, data, radius=None, radius_stride=1):
        if hasattr(self.engine, "route"):
            return self.engine.route(kind, data, self._rank_boxes, self.rank, radius, radius_stride)
        if radius is not None:
            data = torch.cat([data, radius.view(-1)[::radius_stride][:data.shape[0]].unsqueeze(1)], 1)
        return route_generic(kind, data, self._rank_boxes, self.rank)

    def _exchange_remote(self, data, qid_s, send_counts, what, engine=None):
        """Forward the predicates qid_s (grouped by destination), query there, bring the results back.
        -> (query ids [M] (sorted), values [M, 2] (index, rank), distances [M] or None)"""
        engine = engine or self.engine
        dev = data.device
        R = self.world
        rows = torch.cat([data[qid_s].contiguous().view(torch.int32), qid_s.to(torch.int32).unsqueeze(1)], 1)
        fwd, recv_counts = _alltoallv(self.comm, rows, send_counts)
        stride = data.shape[1]
        fwd_preds = fwd[:, :stride].contiguous().view(torch.float32)
        fwd_ids = fwd[:, stride]
        if what[0] == "spatial":
            idx, loff = engine.spatial(self._bottom, what[1], fwd_preds)
            cols = [idx.to(torch.int32).unsqueeze(1)]
        else:
            idx, loff, d = engine.nearest(self._bottom, fwd_preds, what[1])
            cols = [idx.to(torch.int32).unsqueeze(1), d.contiguous().view(torch.int32).unsqueeze(1)]
        loff = loff.long()
        # output_size: known from the result's shape, spares the host sync of a data-dependent size
        res_ids = torch.repeat_interleave(fwd_ids, loff[1:] - loff[:-1], output_size=int(idx.shape[0]))
        back_rows = torch.cat(cols + [res_ids.unsqueeze(1)], 1)
        seg = torch.cumsum(torch.tensor([0] + recv_counts, device=dev), 0)
        back_counts = (loff[seg[1:]] - loff[seg[:-1]]).tolist()
        got, got_counts = _alltoallv(self.comm, back_rows, back_counts)
        src_rank = torch.repeat_interleave(torch.arange(R, device=dev, dtype=torch.int32),
                                           torch.tensor(got_counts, device=dev), output_size=int(got.shape[0]))
        ids = got[:, -1].long()
        order = torch.argsort(ids, stable=True)
        vals = torch.stack([got[:, 0][order], src_rank[order]], 1)
        dists = got[:, 1][order].contiguous().view(torch.float32) if what[0] == "nearest" else None
        return ids[order], vals, dists

    def _spatial_fast(self, kind, data):
        dev = data.device
        q = data.shape[0]
        _mark("start")
        qid_s, send_counts = self._route(kind, data)
        _mark("route")
        if self._overlap and isinstance(self.engine, CudaEngine):
            # the exchange (forward, remote queries for other ranks, results back) runs on a
            # high-priority side stream while a helper thread drives the big local query
            # (abx_query_spatial_crs blocks its host thread once for nnz)
            import threading
            main = self.space.stream
            side = self._side_space()
            side.stream.wait_stream(main)
            box = {}

            def local():
                try:
                    torch.cuda.set_device(data.device)
                    box["l"] = self.engine.spatial(self._bottom, kind, data)
                except BaseException as e:  # re-raised on the calling thread
                    box["e"] = e

            th = threading.Thread(target=local)
            th.start()
            with torch.cuda.stream(side.stream):
                ids, rvals, _ = self._exchange_remote(data, qid_s, send_counts, ("spatial", kind), self._side_engine)
            th.join()
            if "e" in box:
                raise box["e"]
            idx_l, off_l = box["l"]
            main.wait_stream(side.stream)
            ids.record_stream(main)
            rvals.record_stream(main)
        else:
            idx_l, off_l = self.engine.spatial(self._bottom, kind, data)
            _mark("local")
            ids, rvals, _ = self._exchange_remote(data, qid_s, send_counts, ("spatial", kind))
        _mark("exchange")
        if ids.shape[0] == 0:
            # nothing came back from other ranks: the local CRS is the answer
            if hasattr(self.engine, "pair_with_rank"):
                return self.engine.pair_with_rank(idx_l, self.rank), off_l.to(torch.int32)
            return torch.stack([idx_l.to(torch.int32), torch.full_like(idx_l, self.rank, dtype=torch.int32)], 1), \
                off_l.to(torch.int32)
        if hasattr(self.engine, "merge_sorted"):
            _mark("roff")
            out = self.engine.merge_sorted(off_l, idx_l, self.rank, ids, rvals)
        else:
            roff = torch.zeros(q + 1, dtype=torch.int64, device=dev)
            roff[1:] = torch.cumsum(torch.bincount(ids, minlength=q), 0)
            _mark("roff")
            out = self.engine.merge_rows(off_l, idx_l, self.rank, roff, rvals)
        _mark("merge")
        _report("spatial")
        return out

    def _nearest_fast(self, pts, k):
        """Every rank holds >= k primitives: the local k-th distance bounds the true one (phase I
        without an exchange), ranks within that distance are asked for their k nearest (phase II),
        and only the queries that received remote candidates are re-ranked."""
        dev = pts.device
        q = pts.shape[0]
        _mark("start")
        vals = None
        if hasattr(self.engine, "nearest_pairs"):
            # local rows written directly as (index, rank) pairs
            got = self.engine.nearest_pairs(self._bottom, pts, k, self.rank)
            if got is None or got[0].shape[0] != q * k:
                return self._nearest(pts, k)  # short local rows (fewer than k leaves, unreachable leaves)
            vals, d_l = got
            _mark("local")
        else:
            idx_l, off_l, d_l = self.engine.nearest(self._bottom, pts, k)
            _mark("local")
            if idx_l.shape[0] != q * k:
                return self._nearest(pts, k)  # short local rows (unreachable leaves): generic path
        # phase II spheres (point, local k-th distance): routed without materialising them
        qid_s, send_counts = self._route(SPHERE_PRED, pts, d_l.view(-1)[k - 1:], k)
        _mark("route")
        ids, rvals, rd = self._exchange_remote(pts, qid_s, send_counts, ("nearest", k))
        _mark("exchange")
        if vals is None:
            if hasattr(self.engine, "pair_with_rank"):
                vals = self.engine.pair_with_rank(idx_l, self.rank)
            else:
                vals = torch.stack([idx_l.to(torch.int32), torch.full_like(idx_l, self.rank, dtype=torch.int32)], 1)
        _mark("pair")
        out_d = d_l
        if ids.shape[0] and hasattr(self.engine, "knn_merge"):
            self.engine.knn_merge(ids, rvals, rd, k, vals, out_d)
        elif ids.shape[0]:
            # queries with remote candidates: k local + m remote, keep the k smallest
            uq, inv = torch.unique(ids, return_inverse=True)
            cnt = torch.bincount(inv, minlength=uq.shape[0])
            m = int(cnt.max())
            start = torch.cumsum(cnt, 0) - cnt
            pos = torch.arange(ids.shape[0], device=dev) - start[inv]
            u = uq.shape[0]
            cd = torch.full((u, k + m), float("inf"), dtype=torch.float32, device=dev)
            cv = torch.zeros((u, k + m, 2), dtype=torch.int32, device=dev)
            rows = (uq.unsqueeze(1) * k + torch.arange(k, device=dev)).view(-1)  # local rows of the affected queries
            cd[:, :k] = out_d[rows].view(u, k)
            cv[:, :k] = vals[rows].view(u, k, 2)
            cd[inv, k + pos] = rd
            cv[inv, k + pos] = rvals
            sd, so = torch.sort(cd, dim=1, stable=True)
            so = so[:, :k]
            out_d[rows] = sd[:, :k].reshape(-1)
            vals[rows] = torch.gather(cv, 1, so.unsqueeze(2).expand(u, k, 2)).reshape(-1, 2)
        offsets = torch.arange(q + 1, device=dev, dtype=torch.int32) * k
        _mark("rerank")
        _report("nearest")
        return vals, offsets, out_d

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
