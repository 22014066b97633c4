"""ctypes binding of libabx.so (include/abx.h).  No CPU fallback: if the shared
library is missing the import fails, and every compute entry point fails with
ABX_ERR_CUDA when there is no CUDA device."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ABX_LIBRARY: alternative build of the same library (tuning experiments), still no fallback
LIB_PATH = os.environ.get("ABX_LIBRARY") or os.path.join(_HERE, "lib", "libabx.so")

ABX_OK, ABX_ERR_SEARCH, ABX_ERR_CUDA, ABX_ERR_PRECISION, ABX_ERR_ARG = 0, 1, 2, 3, 4


class SearchException(Exception):
    """Mirror of ArborX::SearchException (misc/ArborX_Exception.hpp:19-38)."""


class AbxPolicy(C.Structure):
    _fields_ = [("buffer_size", C.c_int32), ("sort_predicates", C.c_int32)]


ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_size_t)

_vp, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_pp = C.POINTER(C.c_void_p)
_pol = C.POINTER(AbxPolicy)
_pi64 = C.POINTER(C.c_int64)

SIGNATURES = {
    "abx_last_error": (C.c_char_p, []),
    "abx_version": (C.c_int, []),
    "abx_launch_count": (_i64, []),
    "abx_free": (C.c_int, [_vp, _vp]),
    "abx_trim": (_i64, []),
    "abx_profile_enable": (C.c_int, [C.c_int]),
    "abx_profile_report": (_i64, [C.c_char_p, _i64]),
    "abx_bvh_build": (C.c_int, [_vp, C.c_int, _vp, _i64, _pp]),
    "abx_bvh_build_host": (C.c_int, [_vp, C.c_int, _vp, _i64, _pp]),
    "abx_bvh_build_indexed_triangles": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _pp]),
    "abx_bvh_build_from_sorted_codes": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, _pp]),
    "abx_bvh_destroy": (C.c_int, [_vp]),
    "abx_bvh_size": (_i64, [_vp]),
    "abx_bvh_empty": (C.c_int, [_vp]),
    "abx_bvh_bounds": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "abx_bvh_memory_bytes": (_i64, [_vp]),
    "abx_bvh_export_reference_layout": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "abx_query_spatial_crs": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _pol, ALLOC_FN, _vp, _pp, _pp, _pi64]),
    "abx_query_spatial_count": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, C.c_int, _i32, _vp]),
    "abx_query_nearest_crs": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _pol, ALLOC_FN, _vp, _pp, _pp, _pp, _pi64]),
    "abx_query_nearest_geom_crs": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _i32, _pol, ALLOC_FN, _vp, _pp, _pp, _pp,
                                             _pi64]),
    "abx_query_spatial_crs_host": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _pol, ALLOC_FN, _vp, _pp, _pp, _pi64]),
    "abx_query_nearest_crs_host": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _pol, ALLOC_FN, _vp, _pp, _pp, _pp, _pi64]),
    "abx_half_traversal_pairs": (C.c_int, [_vp, _vp, _f, _vp, _i64, _pi64]),
    "abx_find_half_neighbor_list": (C.c_int, [_vp, _vp, _i64, _f, ALLOC_FN, _vp, _pp, _pp, _pi64]),
    "abx_find_full_neighbor_list": (C.c_int, [_vp, _vp, _i64, _f, ALLOC_FN, _vp, _pp, _pp, _pi64]),
    "abx_brute_create": (C.c_int, [_vp, C.c_int, _vp, _i64, _pp]),
    "abx_brute_destroy": (C.c_int, [_vp]),
    "abx_brute_size": (_i64, [_vp]),
    "abx_brute_bounds": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "abx_brute_query_spatial_crs": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, ALLOC_FN, _vp, _pp, _pp, _pi64]),
    "abx_brute_query_nearest_crs": (C.c_int, [_vp, _vp, _vp, _i64, _i32, ALLOC_FN, _vp, _pp, _pp, _pp, _pi64]),
    "abx_dbscan": (C.c_int, [_vp, _vp, _i64, _f, _i32, C.c_int, C.c_int, _vp]),
    "abx_dbscan_host": (C.c_int, [_vp, _vp, _i64, _f, _i32, C.c_int, C.c_int, _vp]),
    "abx_mst_points3f": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "abx_mst_points3f_host": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "abx_dendrogram_union_find": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "abx_hdbscan_points3f": (C.c_int, [_vp, _vp, _i64, _i32, C.c_int, _vp, _vp]),
    "abx_mst_hdbscan_points3f": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "abx_dist_merge_crs": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "abx_bvh_device_view": (C.c_int, [_vp, _vp]),
    "abx_dist_merge_sorted": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "abx_dist_route_count": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _i64, _vp, _i32, _i32, _vp]),
    "abx_dist_route_fill": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "abx_dist_pair_with_rank": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "abx_dist_nearest_pairs": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, C.POINTER(_i64)]),
    "abx_dist_knn_merge": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp]),
    "abx_comm_from_nccl": (C.c_int, [_vp, _pp]),
    "abx_comm_unique_id": (C.c_int, [C.c_char_p]),
    "abx_comm_init_rank": (C.c_int, [C.c_char_p, _i32, _i32, _pp]),
    "abx_comm_create_local": (C.c_int, [_i32, _pp]),
    "abx_comm_destroy": (C.c_int, [_vp]),
    "abx_comm_rank": (_i32, [_vp]),
    "abx_comm_size": (_i32, [_vp]),
    "abx_dist_create": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _pp]),
    "abx_dist_create_host": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _pp]),
    "abx_dist_destroy": (C.c_int, [_vp]),
    "abx_dist_size": (_i64, [_vp]),
    "abx_dist_empty": (C.c_int, [_vp]),
    "abx_dist_bounds": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "abx_dist_query_spatial_crs": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, ALLOC_FN, _vp, _pp, _pp, _pi64]),
    "abx_dist_query_nearest_crs": (C.c_int, [_vp, _vp, _vp, _i64, _i32, ALLOC_FN, _vp, _pp, _pp, _pp, _pi64]),
    "abx_dist_dbscan_points3f": (C.c_int, [_vp, _vp, _vp, _i64, _f, _i32, C.c_int, C.c_int, _vp]),
    "abx_dist_query_spatial_crs_host": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, ALLOC_FN, _vp, _pp, _pp, _pi64, _pp,
                                                  _pp, _pi64]),
    "abx_dist_query_nearest_crs_host": (C.c_int, [_vp, _vp, _vp, _i64, _i32, ALLOC_FN, _vp, _pp, _pp, _pp, _pi64, _pp,
                                                  _pp, _pi64]),
    "abx_scene_bounds": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp]),
    "abx_morton64": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _vp]),
    "abx_morton32": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _vp]),
    "abx_sort_u64": (C.c_int, [_vp, _vp, _vp, _i64]),
    "abx_sort_u32": (C.c_int, [_vp, _vp, _vp, _i64]),
    "abx_exclusive_scan_i32": (C.c_int, [_vp, _vp, _vp, _i64]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "arborx_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C arborx_b200/csrc).  There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)  # AttributeError here = missing export
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status == ABX_OK:
        return
    msg = lib().abx_last_error().decode("utf-8", "replace")
    if status == ABX_ERR_SEARCH:
        raise SearchException(msg)
    if status == ABX_ERR_PRECISION:
        raise RuntimeError(msg)
    if status == ABX_ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError("libabx: CUDA failure: " + msg)
