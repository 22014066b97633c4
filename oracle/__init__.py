"""ctypes loader for the CPU oracle (oracle/arborx_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(arborx_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborc.so")

PRIM_POINT, PRIM_BOX, PRIM_TRI = 0, 1, 2
PRED_SPHERE, PRED_BOX, PRED_POINT, PRED_RAY = 0, 1, 2, 3
PRIM_STRIDE = {PRIM_POINT: 3, PRIM_BOX: 6, PRIM_TRI: 9}
PRED_STRIDE = {PRED_SPHERE: 4, PRED_BOX: 6, PRED_POINT: 3, PRED_RAY: 6}


def build(force=False):
    src = os.path.join(_HERE, "arborx_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "all"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int)
    up = C.POINTER(C.c_uint)
    u64p = C.POINTER(C.c_ulonglong)
    llp = C.POINTER(C.c_longlong)
    vp = C.c_void_p
    sig = {
        "orc_num_threads": (C.c_int, []),
        "orc_set_num_threads": (None, [C.c_int]),
        "orc_expand_bits2_32": (C.c_uint, [C.c_uint]),
        "orc_expand_bits2_64": (C.c_ulonglong, [C.c_ulonglong]),
        "orc_morton32": (C.c_uint, [C.c_float] * 3),
        "orc_morton64": (C.c_ulonglong, [C.c_float] * 3),
        "orc_distance_point_point": (C.c_float, [fp, fp]),
        "orc_distance_point_box": (C.c_float, [fp, fp]),
        "orc_distance_point_triangle": (C.c_float, [fp, fp]),
        "orc_closest_point_triangle": (None, [fp, fp, fp]),
        "orc_intersects": (C.c_int, [C.c_int, fp, C.c_int, fp]),
        "orc_scene_bounds": (None, [C.c_int, fp, C.c_int, fp]),
        "orc_morton64_codes": (None, [C.c_int, fp, C.c_int, fp, u64p]),
        "orc_morton32_codes": (None, [C.c_int, fp, C.c_int, fp, up]),
        "orc_sort_u64": (None, [u64p, C.c_int, up]),
        "orc_bvh_build": (vp, [C.c_int, fp, C.c_int]),
        "orc_bvh_from_sorted_codes": (vp, [C.c_int, fp, u64p, C.c_int]),
        "orc_bvh_destroy": (None, [vp]),
        "orc_bvh_size": (C.c_int, [vp]),
        "orc_bvh_bounds": (None, [vp, fp]),
        "orc_bvh_export": (None, [vp, ip, up, ip, ip, fp, u64p]),
        "orc_query_spatial_count": (None, [vp, C.c_int, fp, C.c_int, C.c_int, ip, llp]),
        "orc_query_spatial_crs": (C.c_longlong, [vp, C.c_int, fp, C.c_int, C.c_int, C.c_int, ip, up]),
        "orc_query_nearest_crs": (C.c_longlong, [vp, fp, C.c_int, C.c_int, ip, C.c_int, ip, up, fp, llp]),
        "orc_query_nearest_geom_crs": (C.c_longlong, [vp, C.c_int, fp, C.c_int, C.c_int, C.c_int, ip, up, fp]),
        "orc_query_ordered_ray_crs": (C.c_longlong, [vp, fp, C.c_int, C.c_int, ip, up, fp]),
        "orc_half_traversal_pairs": (C.c_longlong, [vp, C.c_float, up, C.c_longlong]),
        "orc_union_find_merge": (None, [ip, C.c_int, C.c_int]),
        "orc_union_find_merge_into": (None, [ip, C.c_int, C.c_int]),
        "orc_union_find_representative": (C.c_int, [ip, C.c_int]),
        "orc_dbscan": (C.c_int, [fp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, ip, ip, llp]),
        "orc_dbscan_verify": (C.c_int, [fp, C.c_int, C.c_float, C.c_int, ip, C.c_int]),
        "orc_mst": (C.c_int, [fp, C.c_int, C.c_int, ip, fp]),
        "orc_dendrogram_union_find": (None, [ip, fp, C.c_int, ip, fp]),
        "orc_mst_hdbscan": (C.c_int, [fp, C.c_int, C.c_int, ip, fp, ip, fp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def num_threads():
    return lib().orc_num_threads()


class Tree:
    """Reference-layout BVH built by the oracle (ArborX_LinearBVH.hpp:171-256)."""

    def __init__(self, prims, kind=PRIM_POINT, sorted_codes=None):
        prims = _f32(prims).reshape(-1, PRIM_STRIDE[kind])
        self.kind = kind
        self.n = prims.shape[0]
        self._prims = prims
        if sorted_codes is None:
            self._h = lib().orc_bvh_build(kind, _p(prims, C.c_float), self.n)
        else:
            codes = np.ascontiguousarray(sorted_codes, dtype=np.uint64)
            self._h = lib().orc_bvh_from_sorted_codes(kind, _p(prims, C.c_float), _p(codes, C.c_ulonglong), self.n)

    def __del__(self):
        # at interpreter shutdown the module globals (lib) may already be gone
        if getattr(self, "_h", None) and callable(lib):
            lib().orc_bvh_destroy(self._h)
            self._h = None

    def bounds(self):
        out = np.empty(6, np.float32)
        lib().orc_bvh_bounds(self._h, _p(out, C.c_float))
        return out

    def export(self):
        n = self.n
        m = max(n - 1, 0)
        d = dict(leaf_rope=np.empty(n, np.int32), leaf_index=np.empty(n, np.uint32),
                 left_child=np.empty(m, np.int32), rope=np.empty(m, np.int32),
                 boxes=np.empty((m, 6), np.float32), codes=np.empty(n, np.uint64))
        lib().orc_bvh_export(self._h, _p(d["leaf_rope"], C.c_int), _p(d["leaf_index"], C.c_uint),
                             _p(d["left_child"], C.c_int), _p(d["rope"], C.c_int), _p(d["boxes"], C.c_float),
                             _p(d["codes"], C.c_ulonglong))
        return d

    def spatial_count(self, preds, kind=PRED_SPHERE, limit=0, counters=False):
        preds = _f32(preds).reshape(-1, PRED_STRIDE[kind])
        q = preds.shape[0]
        counts = np.empty(q, np.int32)
        ctr = np.zeros(2, np.int64) if counters else None
        lib().orc_query_spatial_count(self._h, kind, _p(preds, C.c_float), q, limit, _p(counts, C.c_int),
                                      _p(ctr, C.c_longlong))
        return (counts, ctr) if counters else counts

    def spatial_crs(self, preds, kind=PRED_SPHERE, sort_predicates=True, buffer_size=0):
        preds = _f32(preds).reshape(-1, PRED_STRIDE[kind])
        q = preds.shape[0]
        offsets = np.zeros(q + 1, np.int32)
        nnz = lib().orc_query_spatial_crs(self._h, kind, _p(preds, C.c_float), q, int(sort_predicates), buffer_size,
                                          _p(offsets, C.c_int), None)
        if nnz < 0:
            raise RuntimeError("SearchException: hard preallocation overflow")
        indices = np.empty(nnz, np.uint32)
        if nnz:
            lib().orc_query_spatial_crs(self._h, kind, _p(preds, C.c_float), q, int(sort_predicates), buffer_size,
                                        _p(offsets, C.c_int), _p(indices, C.c_uint))
        return offsets, indices

    def nearest_crs(self, pts, k, sort_predicates=True, counters=False):
        pts = _f32(pts).reshape(-1, 3)
        q = pts.shape[0]
        kk = None
        if np.ndim(k) > 0:
            kk = np.ascontiguousarray(k, dtype=np.int32)
            total = int(np.maximum(kk, 0).sum())
            k0 = 0
        else:
            total = q * max(int(k), 0)
            k0 = int(k)
        offsets = np.zeros(q + 1, np.int32)
        indices = np.empty(total, np.uint32)
        dists = np.empty(total, np.float32)
        ctr = np.zeros(2, np.int64) if counters else None
        nnz = lib().orc_query_nearest_crs(self._h, _p(pts, C.c_float), q, k0, _p(kk, C.c_int), int(sort_predicates),
                                          _p(offsets, C.c_int), _p(indices, C.c_uint), _p(dists, C.c_float),
                                          _p(ctr, C.c_longlong))
        res = (offsets, indices[:nnz], dists[:nnz])
        return res + (ctr,) if counters else res

    def nearest_geom_crs(self, preds, kind, k, sort_predicates=True):
        """nearest(Box | Sphere | Ray | Point, k) -> (offsets, indices, distances)."""
        preds = _f32(preds).reshape(-1, PRED_STRIDE[kind])
        q = preds.shape[0]
        offsets = np.zeros(q + 1, np.int32)
        indices = np.empty(q * max(int(k), 0), np.uint32)
        dists = np.empty(q * max(int(k), 0), np.float32)
        nnz = lib().orc_query_nearest_geom_crs(self._h, kind, _p(preds, C.c_float), q, int(k), int(sort_predicates),
                                               _p(offsets, C.c_int), _p(indices, C.c_uint), _p(dists, C.c_float))
        if nnz < 0:
            raise ValueError("nearest(geometry, k): point and box primitives only")
        return offsets, indices[:nnz], dists[:nnz]

    def ordered_ray_crs(self, rays, limit=0):
        """ordered_intersects(ray): per ray the leaves in the order the reference's traversal hands them to the
        callback (limit > 0: the callback exits after `limit` calls) -> (offsets, indices, distances)."""
        rays = _f32(rays).reshape(-1, 6)
        q = rays.shape[0]
        offsets = np.zeros(q + 1, np.int32)
        nnz = lib().orc_query_ordered_ray_crs(self._h, _p(rays, C.c_float), q, int(limit), _p(offsets, C.c_int), None, None)
        if nnz < 0:
            raise ValueError("ordered_intersects(ray): point and box primitives only")
        indices = np.empty(nnz, np.uint32)
        dists = np.empty(nnz, np.float32)
        lib().orc_query_ordered_ray_crs(self._h, _p(rays, C.c_float), q, int(limit), _p(offsets, C.c_int),
                                        _p(indices, C.c_uint), _p(dists, C.c_float))
        return offsets, indices, dists

    def half_pairs(self, r):
        cnt = lib().orc_half_traversal_pairs(self._h, C.c_float(r), None, 0)
        pairs = np.empty((cnt, 2), np.uint32)
        lib().orc_half_traversal_pairs(self._h, C.c_float(r), _p(pairs, C.c_uint), cnt)
        return pairs


def scene_bounds(prims, kind=PRIM_POINT):
    prims = _f32(prims).reshape(-1, PRIM_STRIDE[kind])
    out = np.empty(6, np.float32)
    lib().orc_scene_bounds(kind, _p(prims, C.c_float), prims.shape[0], _p(out, C.c_float))
    return out


def morton64_codes(prims, bounds6, kind=PRIM_POINT):
    prims = _f32(prims).reshape(-1, PRIM_STRIDE[kind])
    b = _f32(bounds6)
    out = np.empty(prims.shape[0], np.uint64)
    lib().orc_morton64_codes(kind, _p(prims, C.c_float), prims.shape[0], _p(b, C.c_float), _p(out, C.c_ulonglong))
    return out


def morton32_codes(preds, bounds6, kind=PRED_SPHERE):
    preds = _f32(preds).reshape(-1, PRED_STRIDE[kind])
    b = _f32(bounds6)
    out = np.empty(preds.shape[0], np.uint32)
    lib().orc_morton32_codes(kind, _p(preds, C.c_float), preds.shape[0], _p(b, C.c_float), _p(out, C.c_uint))
    return out


def sort_u64(keys):
    keys = np.array(keys, dtype=np.uint64)
    perm = np.empty(keys.shape[0], np.uint32)
    lib().orc_sort_u64(_p(keys, C.c_ulonglong), keys.shape[0], _p(perm, C.c_uint))
    return keys, perm


def dbscan(xyz, eps, minpts, impl=0, algo=0, return_core=False, return_stats=False):
    """impl: 0 FDBSCAN, 1 FDBSCAN_DenseBox; algo: 0 DBSCAN, 1 DBSCAN* (ArborX_DBSCAN.hpp:180-216)."""
    xyz = _f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    labels = np.empty(n, np.int32)
    core = np.empty(n, np.int32)
    stats = np.zeros(4, np.int64)
    rc = lib().orc_dbscan(_p(xyz, C.c_float), n, C.c_float(eps), minpts, impl, algo, _p(labels, C.c_int),
                          _p(core, C.c_int), _p(stats, C.c_longlong))
    if rc == 1:
        raise ValueError("SearchException: eps > 0 and minpts >= 2 required")
    if rc == 2:
        raise RuntimeError("FDBSCAN-DenseBox loss of precision")
    out = (labels,)
    if return_core:
        out += (core.astype(bool),)
    if return_stats:
        out += (stats,)
    return out if len(out) > 1 else labels


def dbscan_verify(xyz, eps, minpts, labels, algo=0):
    """Returns the bit mask of failed verifier checks (0 = accepted)."""
    xyz = _f32(xyz).reshape(-1, 3)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    return lib().orc_dbscan_verify(_p(xyz, C.c_float), xyz.shape[0], C.c_float(eps), minpts, _p(labels, C.c_int), algo)


def mst(xyz, k=1, return_iterations=False):
    """Euclidean (k = 1) or mutual-reachability (k > 1) minimum spanning tree
    (ArborX_MinimumSpanningTree.hpp:46-101) -> (edges int32 [n - 1, 2] in original indices, weights float32)."""
    xyz = _f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    edges = np.empty((max(n - 1, 0), 2), np.int32)
    weights = np.empty(max(n - 1, 0), np.float32)
    it = lib().orc_mst(_p(xyz, C.c_float), n, int(k), _p(edges, C.c_int), _p(weights, C.c_float))
    return (edges, weights, it) if return_iterations else (edges, weights)


def dendrogram(edges, weights):
    """Dendrogram(space, edges), UNION_FIND (ArborX_Dendrogram.hpp:47-76) -> (parents [2 e + 1], parent_heights [e])."""
    edges = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 2)
    weights = _f32(weights).reshape(-1)
    e = edges.shape[0]
    parents = np.full(2 * e + 1, -1, np.int32)
    heights = np.empty(e, np.float32)
    lib().orc_dendrogram_union_find(_p(edges, C.c_int), _p(weights, C.c_float), e, _p(parents, C.c_int),
                                    _p(heights, C.c_float))
    return parents, heights


def mst_hdbscan(xyz, k=1):
    """MinimumSpanningTree<..., BoruvkaMode::HDBSCAN> (ArborX_MinimumSpanningTree.hpp:31-297) -> (edges, weights,
    dendrogram_parents [2 n - 1], dendrogram_parent_heights [n - 1]); the edges are in the hybrid algorithm's own
    (chain, weight) order and the parents index that order."""
    xyz = _f32(xyz).reshape(-1, 3)
    n = xyz.shape[0]
    edges = np.empty((max(n - 1, 0), 2), np.int32)
    weights = np.empty(max(n - 1, 0), np.float32)
    parents = np.full(max(2 * n - 1, 0), -1, np.int32)
    heights = np.empty(max(n - 1, 0), np.float32)
    lib().orc_mst_hdbscan(_p(xyz, C.c_float), n, int(k), _p(edges, C.c_int), _p(weights, C.c_float),
                          _p(parents, C.c_int), _p(heights, C.c_float))
    return edges, weights, parents, heights
