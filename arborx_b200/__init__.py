"""arborx_b200 -- host-side mirror of the ArborX interface for the hot path
(BoundingVolumeHierarchy, query with intersects/nearest predicates, dbscan) over
the C ABI of libabx.so (include/abx.h).  torch is used only for device memory and
streams.  Spellings follow the reference:

    bvh = BoundingVolumeHierarchy(space, points)            # ArborX_LinearBVH.hpp:68-73
    indices, offsets = bvh.query(space, make_intersects(q, r))   # :90-110
    indices, offsets = bvh.query(space, make_nearest(q, k))
    labels = dbscan(space, points, eps, minpts, DBSCANParameters())  # ArborX_DBSCAN.hpp:219-223
"""
import ctypes as C

import torch

from . import _lib
from ._lib import SearchException, lib

POINT, BOX, TRIANGLE = 0, 1, 2            # primitive kinds (ABX_PRIM_*)
SPHERE_PRED, BOX_PRED, POINT_PRED, RAY_PRED = 0, 1, 2, 3  # predicate geometries (ABX_PRED_*)
_PRIM_STRIDE = {POINT: 3, BOX: 6, TRIANGLE: 9}
_PRED_STRIDE = {SPHERE_PRED: 4, BOX_PRED: 6, POINT_PRED: 3, RAY_PRED: 6}

__all__ = ["ExecutionSpace", "BoundingVolumeHierarchy", "BVH", "TraversalPolicy", "HostBufferPool", "BruteForce", "intersects", "nearest",
           "make_intersects", "make_nearest", "query", "dbscan", "DBSCANParameters", "SearchException",
           "MinimumSpanningTree", "Dendrogram", "hdbscan", "DENDROGRAM_BORUVKA", "DENDROGRAM_UNION_FIND",
           "POINT", "BOX", "TRIANGLE", "launch_count"]


class ExecutionSpace:
    """Execution space instance = a CUDA stream; every kernel of a call is enqueued on it
    (reference: the `space` argument of every API, tstKokkosToolsExecutionSpaceInstances.cpp:80-153)."""

    def __init__(self, stream=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("arborx_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)

    @property
    def handle(self):
        return C.c_void_p(self.stream.cuda_stream)

    def fence(self):
        self.stream.synchronize()


class TraversalPolicy:
    """Experimental::TraversalPolicy (detail/ArborX_TraversalPolicy.hpp:19-48)."""

    def __init__(self, buffer_size=0, sort_predicates=True):
        self._buffer_size = buffer_size
        self._sort_predicates = sort_predicates

    def setBufferSize(self, b):
        self._buffer_size = b
        return self

    def setPredicateSorting(self, s):
        self._sort_predicates = s
        return self

    def _c(self):
        return _lib.AbxPolicy(int(self._buffer_size), int(bool(self._sort_predicates)))


class Predicates:
    def __init__(self, tag, kind, data, k=None):
        self.tag = tag      # "spatial" | "nearest"
        self.kind = kind
        self.data = data    # float32 [q, stride], device or host tensor
        self.k = k

    def size(self):
        return self.data.shape[0]


def _as_f32(t, stride):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t, dtype=torch.float32)
    t = t.to(torch.float32).reshape(-1, stride).contiguous()
    return t


def intersects(geometry, kind=None):
    """intersects(Sphere|Box|Point|Ray) for a batch (detail/ArborX_Predicates.hpp:130-147):
    [q,4] spheres (centre, radius), [q,6] boxes or [q,3] points; rays ([q,6] origin, direction,
    geometry/ArborX_Ray.hpp) need kind=RAY_PRED."""
    if not isinstance(geometry, torch.Tensor):
        geometry = torch.as_tensor(geometry, dtype=torch.float32)
    if kind is None:
        kind = {4: SPHERE_PRED, 6: BOX_PRED, 3: POINT_PRED}[geometry.shape[-1]]
    return Predicates("spatial", kind, _as_f32(geometry, _PRED_STRIDE[kind]))


def nearest(geometry, k=1, kind=None):
    """nearest(Geometry, k) for a batch (detail/ArborX_Predicates.hpp:58-80,130-135): [q, 3] points (k may be a
    tensor of per-query k), [q, 6] boxes, [q, 4] spheres; rays ([q, 6] origin, direction) need kind=RAY_PRED."""
    if not isinstance(geometry, torch.Tensor):
        geometry = torch.as_tensor(geometry, dtype=torch.float32)
    if kind is None:
        kind = {3: POINT_PRED, 6: BOX_PRED, 4: SPHERE_PRED}[geometry.shape[-1]]
    return Predicates("nearest", kind, _as_f32(geometry, _PRED_STRIDE[kind]), k)


def make_intersects(points, r):
    """Experimental::make_intersects(points, r): spheres of radius r (detail/ArborX_PredicateHelpers.hpp:93-105)."""
    p = _as_f32(points, 3)
    rr = torch.full((p.shape[0], 1), float(r), dtype=torch.float32, device=p.device)
    return Predicates("spatial", SPHERE_PRED, torch.cat([p, rr], 1).contiguous())


def make_nearest(points, k):
    """Experimental::make_nearest(points, k) (PredicateHelpers.hpp:107-119)."""
    return nearest(points, k)


class HostBufferPool:
    """Caller-owned pool of pinned host buffers for the results of host-predicate queries
    (`query(..., out=pool)`).  Results of a query ALIAS the pool's buffers and stay valid until the next
    query that is given the same pool; without a pool every host query returns freshly allocated pinned
    tensors that the caller owns.  (cudaHostAlloc of a few hundred MB per query dominates a short
    end-to-end step, which is why bench.py passes one pool per host thread.)"""

    def __init__(self):
        self._bufs = {}

    def take(self, which, n, dtype):
        buf = self._bufs.get(which)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = torch.empty(max(int(n * 1.05) + 16, 1), dtype=dtype, pin_memory=True)
            self._bufs[which] = buf
        return buf[:n]


class _Allocator:
    """abx_alloc_fn backed by torch tensors (the library 'resizes the caller's views')."""
    _DT = {0: torch.int32, 1: torch.int32, 2: torch.float32, 3: torch.int32, 4: torch.int32}

    def __init__(self, device, pinned_host=False, pool=None):
        self.device = device
        self.pinned_host = pinned_host
        self.pool = pool
        self.out = {}
        self.fn = _lib.ALLOC_FN(self._alloc)

    def _alloc(self, user, which, nbytes):
        n = nbytes // 4
        if self.pinned_host:
            if self.pool is not None:
                t = self.pool.take(which, n, self._DT[which])
            else:
                t = torch.empty(n, dtype=self._DT[which], pin_memory=n > 0)
        else:
            t = torch.empty(n, dtype=self._DT[which], device=self.device)
        self.out[which] = t
        return t.data_ptr() if n else None


class BoundingVolumeHierarchy:
    """ArborX::BoundingVolumeHierarchy (spatial/ArborX_LinearBVH.hpp:50-142): leaf value = index of the primitive."""

    def __init__(self, space, values, kind=None):
        if not isinstance(values, torch.Tensor):
            values = torch.as_tensor(values, dtype=torch.float32)
        if kind is None:
            kind = {3: POINT, 6: BOX, 9: TRIANGLE}[values.shape[-1]] if values.dim() == 2 else POINT
        self.kind = kind
        self._space = space
        v = _as_f32(values, _PRIM_STRIDE[kind])
        h = C.c_void_p()
        with torch.cuda.stream(space.stream):
            if v.is_cuda:
                _lib.check(lib().abx_bvh_build(space.handle, kind, C.c_void_p(v.data_ptr()), v.shape[0], C.byref(h)))
            else:
                _lib.check(lib().abx_bvh_build_host(space.handle, kind, C.c_void_p(v.data_ptr()), v.shape[0],
                                                    C.byref(h)))
        self._values = v  # keep the input alive until the build has been enqueued and run
        self._h = h

    @classmethod
    def from_indexed_triangles(cls, space, vertices, triangles):
        """Triangles as vertex-index triples ([V, 3] float32 vertices, [T, 3] int32 indices): the access pattern
        of benchmarks/triangulated_surface_distance (triangulated_surface_distance.cpp:34-58)."""
        self = cls.__new__(cls)
        self.kind = TRIANGLE
        self._space = space
        v = _as_f32(vertices, 3).to(space.device)
        t = torch.as_tensor(triangles).to(device=space.device, dtype=torch.int32).reshape(-1, 3).contiguous()
        h = C.c_void_p()
        with torch.cuda.stream(space.stream):
            _lib.check(lib().abx_bvh_build_indexed_triangles(space.handle, C.c_void_p(v.data_ptr()), v.shape[0],
                                                             C.c_void_p(t.data_ptr()), t.shape[0], C.byref(h)))
        self._values = (v, t)
        self._h = h
        return self

    @classmethod
    def _from_sorted_codes(cls, space, values, codes, kind):
        self = cls.__new__(cls)
        self.kind = kind
        self._space = space
        v = _as_f32(values, _PRIM_STRIDE[kind]).to(space.device)
        c = codes.to(space.device).contiguous()
        h = C.c_void_p()
        with torch.cuda.stream(space.stream):
            _lib.check(lib().abx_bvh_build_from_sorted_codes(space.handle, kind, C.c_void_p(v.data_ptr()),
                                                             C.c_void_p(c.data_ptr()), v.shape[0], C.byref(h)))
        space.fence()
        self._values = v
        self._h = h
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().abx_bvh_destroy(h)
            except Exception:
                pass
            self._h = None

    def size(self):
        return lib().abx_bvh_size(self._h)

    def empty(self):
        return bool(lib().abx_bvh_empty(self._h))

    def bounds(self):
        out = (C.c_float * 6)()
        _lib.check(lib().abx_bvh_bounds(self._h, out))
        return torch.tensor(list(out), dtype=torch.float32)

    def memory_bytes(self):
        return lib().abx_bvh_memory_bytes(self._h)

    def query(self, space, predicates, policy=None, return_distances=False, out=None):
        """query(space, predicates, indices, offsets[, policy]) -> (indices, offsets[, distances]).
        Device predicates give device results; host (CPU tensor) predicates run the host-buffer
        entry points and give pinned host results (the end-to-end path).  The returned tensors are owned
        by the caller; `out` (a HostBufferPool, host predicates only) makes them views of the pool's
        reusable pinned buffers instead, valid until the next query given the same pool."""
        pol = (policy or TraversalPolicy())._c()
        d = predicates.data
        q = d.shape[0]
        host = not d.is_cuda
        alloc = _Allocator(space.device, pinned_host=host, pool=out if host else None)
        off, idx, dist = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nnz = C.c_int64()
        L = lib()
        with torch.cuda.stream(space.stream):
            if predicates.tag == "spatial":
                fn = L.abx_query_spatial_crs_host if host else L.abx_query_spatial_crs
                _lib.check(fn(self._h, space.handle, predicates.kind, C.c_void_p(d.data_ptr()), q, C.byref(pol),
                              alloc.fn, None, C.byref(off), C.byref(idx), C.byref(nnz)))
            else:
                k = predicates.k
                want_d = C.byref(dist) if return_distances else None
                if predicates.kind != POINT_PRED:
                    if host or isinstance(k, torch.Tensor):
                        raise ValueError("nearest(Box | Sphere | Ray, k): device predicates, uniform k")
                    _lib.check(L.abx_query_nearest_geom_crs(self._h, space.handle, predicates.kind,
                                                            C.c_void_p(d.data_ptr()), q, int(k), C.byref(pol), alloc.fn,
                                                            None, C.byref(off), C.byref(idx), want_d, C.byref(nnz)))
                elif isinstance(k, torch.Tensor):
                    if host:
                        raise ValueError("per-query k needs device predicates")
                    kk = k.to(device=space.device, dtype=torch.int32).contiguous()
                    _lib.check(L.abx_query_nearest_crs(self._h, space.handle, C.c_void_p(d.data_ptr()), q, 0,
                                                       C.c_void_p(kk.data_ptr()), C.byref(pol), alloc.fn, None,
                                                       C.byref(off), C.byref(idx), want_d, C.byref(nnz)))
                elif host:
                    _lib.check(L.abx_query_nearest_crs_host(self._h, space.handle, C.c_void_p(d.data_ptr()), q, int(k),
                                                            C.byref(pol), alloc.fn, None, C.byref(off), C.byref(idx),
                                                            want_d, C.byref(nnz)))
                else:
                    _lib.check(L.abx_query_nearest_crs(self._h, space.handle, C.c_void_p(d.data_ptr()), q, int(k),
                                                       None, C.byref(pol), alloc.fn, None, C.byref(off), C.byref(idx),
                                                       want_d, C.byref(nnz)))
        dev = "cpu" if host else space.device
        out = alloc.out
        # break the allocator <-> ctypes-callback reference cycle now: otherwise the result
        # tensors stay alive until the cyclic GC runs and device memory keeps growing
        alloc.out, alloc.fn = {}, None
        indices = out.get(1, torch.empty(0, dtype=torch.int32, device=dev))
        offsets = out[0]
        if return_distances:
            return indices, offsets, out.get(2, torch.empty(0, dtype=torch.float32, device=dev))
        return indices, offsets

    def count(self, space, predicates, limit=0, sort_predicates=True):
        """query(space, predicates, callback) with a counting callback (optionally CountUpToN)."""
        d = predicates.data.to(space.device)
        counts = torch.empty(d.shape[0], dtype=torch.int32, device=space.device)
        with torch.cuda.stream(space.stream):
            _lib.check(lib().abx_query_spatial_count(self._h, space.handle, predicates.kind, C.c_void_p(d.data_ptr()),
                                                     d.shape[0], int(sort_predicates), int(limit),
                                                     C.c_void_p(counts.data_ptr())))
        return counts

    def export_reference_layout(self, space):
        n = self.size()
        m = max(n - 1, 0)
        dev = space.device
        out = dict(leaf_rope=torch.empty(n, dtype=torch.int32, device=dev),
                   leaf_index=torch.empty(n, dtype=torch.int32, device=dev),
                   left_child=torch.empty(m, dtype=torch.int32, device=dev),
                   rope=torch.empty(m, dtype=torch.int32, device=dev),
                   boxes=torch.empty((m, 6), dtype=torch.float32, device=dev),
                   codes=torch.empty(n, dtype=torch.int64, device=dev))
        with torch.cuda.stream(space.stream):
            _lib.check(lib().abx_bvh_export_reference_layout(
                self._h, space.handle, *[C.c_void_p(out[k].data_ptr()) for k in
                                         ("leaf_rope", "leaf_index", "left_child", "rope", "boxes", "codes")]))
        space.fence()
        return out

    def half_traversal_pairs(self, space, r):
        cnt = C.c_int64()
        with torch.cuda.stream(space.stream):
            _lib.check(lib().abx_half_traversal_pairs(self._h, space.handle, float(r), None, 0, C.byref(cnt)))
            pairs = torch.empty((cnt.value, 2), dtype=torch.int32, device=space.device)
            _lib.check(lib().abx_half_traversal_pairs(self._h, space.handle, float(r), C.c_void_p(pairs.data_ptr()),
                                                      cnt.value, C.byref(cnt)))
        return pairs


BVH = BoundingVolumeHierarchy


def _neighbor_list(space, points, radius, full):
    pts = _as_f32(points, 3).to(space.device)
    alloc = _Allocator(space.device)
    off, idx, nnz = C.c_void_p(), C.c_void_p(), C.c_int64()
    L = lib()
    fn = L.abx_find_full_neighbor_list if full else L.abx_find_half_neighbor_list
    with torch.cuda.stream(space.stream):
        _lib.check(fn(space.handle, C.c_void_p(pts.data_ptr()), pts.shape[0], float(radius), alloc.fn, None,
                      C.byref(off), C.byref(idx), C.byref(nnz)))
    out = alloc.out
    alloc.out, alloc.fn = {}, None
    return out[0], out.get(1, torch.empty(0, dtype=torch.int32, device=space.device))


def find_half_neighbor_list(space, points, radius):
    """Experimental::findHalfNeighborList (spatial/detail/ArborX_NeighborList.hpp:47-110): every unordered pair of
    points within `radius` appears once, in the row of the point the half traversal reports second.
    -> (offsets, indices); the order inside a row is unspecified, as in the reference."""
    return _neighbor_list(space, points, radius, False)


def find_full_neighbor_list(space, points, radius):
    """Experimental::findFullNeighborList (ArborX_NeighborList.hpp:112-190): the symmetric list, every pair in both
    rows (the half list expanded, ArborX_ExpandHalfToFull.hpp:24-72)."""
    return _neighbor_list(space, points, radius, True)


class BruteForce:
    """ArborX::BruteForce (spatial/ArborX_BruteForce.hpp:42-160): the BVH's query interface answered by exhaustive
    tests; point and box primitives."""

    def __init__(self, space, values, kind=None):
        if not isinstance(values, torch.Tensor):
            values = torch.as_tensor(values, dtype=torch.float32)
        if kind is None:
            kind = {3: POINT, 6: BOX}[values.shape[-1]]
        self.kind = kind
        v = _as_f32(values, _PRIM_STRIDE[kind]).to(space.device)
        h = C.c_void_p()
        with torch.cuda.stream(space.stream):
            _lib.check(lib().abx_brute_create(space.handle, kind, C.c_void_p(v.data_ptr()), v.shape[0], C.byref(h)))
        self._values = v
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().abx_brute_destroy(h)
            except Exception:
                pass
            self._h = None

    def size(self):
        return lib().abx_brute_size(self._h)

    def empty(self):
        return self.size() == 0

    def bounds(self):
        out = (C.c_float * 6)()
        _lib.check(lib().abx_brute_bounds(self._h, out))
        return torch.tensor(list(out), dtype=torch.float32)

    def query(self, space, predicates, return_distances=False):
        d = predicates.data.to(space.device)
        q = d.shape[0]
        alloc = _Allocator(space.device)
        off, idx, dist = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nnz = C.c_int64()
        L = lib()
        with torch.cuda.stream(space.stream):
            if predicates.tag == "spatial":
                _lib.check(L.abx_brute_query_spatial_crs(self._h, space.handle, predicates.kind, C.c_void_p(d.data_ptr()),
                                                         q, alloc.fn, None, C.byref(off), C.byref(idx), C.byref(nnz)))
            else:
                if predicates.kind != POINT_PRED:
                    raise ValueError("BruteForce: nearest(Point, k)")
                _lib.check(L.abx_brute_query_nearest_crs(self._h, space.handle, C.c_void_p(d.data_ptr()), q,
                                                         int(predicates.k), alloc.fn, None, C.byref(off), C.byref(idx),
                                                         C.byref(dist) if return_distances else None, C.byref(nnz)))
        out = alloc.out
        alloc.out, alloc.fn = {}, None
        indices = out.get(1, torch.empty(0, dtype=torch.int32, device=space.device))
        if return_distances:
            return indices, out[0], out.get(2, torch.empty(0, dtype=torch.float32, device=space.device))
        return indices, out[0]


def query(tree, space, predicates, policy=None, **kw):
    """ArborX::query free function (spatial/ArborX_CrsGraphWrapper.hpp:22-35)."""
    return tree.query(space, predicates, policy, **kw)


class DBSCANParameters:
    """DBSCAN::Parameters (cluster/ArborX_DBSCAN.hpp:192-216); default implementation FDBSCAN_DenseBox."""
    FDBSCAN, FDBSCAN_DenseBox = 0, 1
    DBSCAN, DBSCAN_STAR = 0, 1

    def __init__(self, implementation=1, algorithm=0, verbose=False):
        self._implementation = implementation
        self._algorithm = algorithm
        self._verbose = verbose

    def setImplementation(self, impl):
        self._implementation = impl
        return self

    def setAlgorithm(self, algo):
        self._algorithm = algo
        return self

    def setVerbosity(self, v):
        self._verbose = v
        return self


def dbscan(space, primitives, eps, core_min_size, parameters=None):
    """ArborX::dbscan(space, primitives, eps, core_min_size, labels, parameters) -> labels (int32)."""
    p = parameters or DBSCANParameters()
    x = _as_f32(primitives, 3)
    n = x.shape[0]
    L = lib()
    with torch.cuda.stream(space.stream):
        if x.is_cuda:
            labels = torch.empty(n, dtype=torch.int32, device=space.device)
            _lib.check(L.abx_dbscan(space.handle, C.c_void_p(x.data_ptr()), n, float(eps), int(core_min_size),
                                    p._implementation, p._algorithm, C.c_void_p(labels.data_ptr())))
        else:
            labels = torch.empty(n, dtype=torch.int32, pin_memory=n > 0)
            _lib.check(L.abx_dbscan_host(space.handle, C.c_void_p(x.data_ptr()), n, float(eps), int(core_min_size),
                                         p._implementation, p._algorithm, C.c_void_p(labels.data_ptr())))
    return labels


DENDROGRAM_BORUVKA, DENDROGRAM_UNION_FIND = 0, 1  # DendrogramImplementation (cluster/ArborX_Dendrogram.hpp:24-28)


class MinimumSpanningTree:
    """ArborX::Experimental::MinimumSpanningTree(space, points, k = 1) (cluster/ArborX_MinimumSpanningTree.hpp:
    31-101): `.edges` int32 [n - 1, 2] (source, target) in the caller's indices and `.weights` float32 [n - 1] --
    Euclidean for k = 1, mutual reachability for k > 1.  Device points give device results, host points pinned
    host results.  The order of the edges is unspecified (sort before comparing, as the reference's tests do).
    mode="hdbscan" is BoruvkaMode::HDBSCAN (device points only): the edges come in the hybrid algorithm's (chain,
    weight) order and `.dendrogram_parents` [2 n - 1] / `.dendrogram_parent_heights` [n - 1] index that order."""

    def __init__(self, space, points, k=1, mode="mst"):
        x = _as_f32(points, 3)
        n = x.shape[0]
        m = max(n - 1, 0)
        it = C.c_int32(0)
        L = lib()
        with torch.cuda.stream(space.stream):
            if mode == "hdbscan":
                assert x.is_cuda and n >= 1
                self.edges = torch.empty((m, 2), dtype=torch.int32, device=space.device)
                self.weights = torch.empty(m, dtype=torch.float32, device=space.device)
                self.dendrogram_parents = torch.empty(2 * n - 1, dtype=torch.int32, device=space.device)
                self.dendrogram_parent_heights = torch.empty(m, dtype=torch.float32, device=space.device)
                _lib.check(L.abx_mst_hdbscan_points3f(space.handle, C.c_void_p(x.data_ptr()), n, int(k),
                                                      C.c_void_p(self.edges.data_ptr()),
                                                      C.c_void_p(self.weights.data_ptr()),
                                                      C.c_void_p(self.dendrogram_parents.data_ptr()),
                                                      C.c_void_p(self.dendrogram_parent_heights.data_ptr()),
                                                      C.byref(it)))
                self.iterations = int(it.value)
                return
            if x.is_cuda:
                self.edges = torch.empty((m, 2), dtype=torch.int32, device=space.device)
                self.weights = torch.empty(m, dtype=torch.float32, device=space.device)
                fn = L.abx_mst_points3f
            else:
                self.edges = torch.empty((m, 2), dtype=torch.int32, pin_memory=m > 0)
                self.weights = torch.empty(m, dtype=torch.float32, pin_memory=m > 0)
                fn = L.abx_mst_points3f_host
            _lib.check(fn(space.handle, C.c_void_p(x.data_ptr()), n, int(k), C.c_void_p(self.edges.data_ptr()),
                          C.c_void_p(self.weights.data_ptr()), C.byref(it)))
        self.iterations = int(it.value)


class Dendrogram:
    """ArborX::Experimental::Dendrogram(space, edges) with DendrogramImplementation::UNION_FIND
    (cluster/ArborX_Dendrogram.hpp:31-76): `_parents` int32 [2 e + 1] (edges in ascending weight order, then the
    vertices; root -> -1) and `_parent_heights` float32 [e]."""

    def __init__(self, space, edges, weights):
        e = edges.to(torch.int32).contiguous().view(-1, 2)
        w = weights.to(torch.float32).contiguous().view(-1)
        assert e.is_cuda and w.is_cuda and e.shape[0] == w.shape[0]
        m = e.shape[0]
        self._parents = torch.empty(2 * m + 1, dtype=torch.int32, device=space.device)
        self._parent_heights = torch.empty(m, dtype=torch.float32, device=space.device)
        with torch.cuda.stream(space.stream):
            _lib.check(lib().abx_dendrogram_union_find(space.handle, C.c_void_p(e.data_ptr()), C.c_void_p(w.data_ptr()),
                                                       m, C.c_void_p(self._parents.data_ptr()),
                                                       C.c_void_p(self._parent_heights.data_ptr())))


def hdbscan(space, primitives, core_min_size, dendrogram_impl=DENDROGRAM_BORUVKA):
    """ArborX::Experimental::hdbscan(space, primitives, core_min_size, dendrogram_impl = BORUVKA)
    (cluster/ArborX_HDBSCAN.hpp:29-53) -> Dendrogram-like object with `_parents`, `_parent_heights` (BORUVKA: edges
    in the hybrid algorithm's order, all on the device; UNION_FIND: edges in ascending weight order)."""
    x = _as_f32(primitives, 3)
    assert x.is_cuda
    n = x.shape[0]

    class _D:
        pass
    d = _D()
    d._parents = torch.empty(2 * n - 1, dtype=torch.int32, device=space.device)
    d._parent_heights = torch.empty(max(n - 1, 0), dtype=torch.float32, device=space.device)
    with torch.cuda.stream(space.stream):
        _lib.check(lib().abx_hdbscan_points3f(space.handle, C.c_void_p(x.data_ptr()), n, int(core_min_size),
                                              int(dendrogram_impl), C.c_void_p(d._parents.data_ptr()),
                                              C.c_void_p(d._parent_heights.data_ptr())))
    return d


def launch_count():
    return lib().abx_launch_count()


def trim():
    """Return the library's cached device blocks (and torch's) to the driver; -> bytes released by libabx."""
    released = lib().abx_trim()
    torch.cuda.empty_cache()
    return released


def profile_enable(on=True):
    """Start/stop per-kernel CUDA-event timing inside the library."""
    _lib.check(lib().abx_profile_enable(int(bool(on))))


def profile_report():
    """-> list of (kernel name, launches, total_ms, max_ms of one launch), most expensive first."""
    need = lib().abx_profile_report(None, 0)
    buf = C.create_string_buffer(int(need) + 16)
    lib().abx_profile_report(buf, len(buf))
    rows = []
    for line in buf.value.decode().splitlines():
        name, cnt, ms, mx = line.rsplit("\t", 3)
        rows.append((name, int(cnt), float(ms), float(mx)))
    return rows
