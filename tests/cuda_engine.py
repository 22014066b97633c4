"""Adapter: the CUDA product (arborx_b200 -> libabx.so C ABI) behind the numpy
engine interface of tests/engines.py."""
import ctypes as C

import numpy as np
import torch

import arborx_b200 as abx
from arborx_b200 import _lib
from tests.engines import PRED_SPHERE, PRIM_POINT, SearchException


def _dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


class _CudaTree:
    def __init__(self, space, bvh, n, kind):
        self.space, self.bvh, self.n, self.kind = space, bvh, n, kind

    def bounds(self):
        return self.bvh.bounds().numpy()

    def export(self):
        d = self.bvh.export_reference_layout(self.space)
        out = {k: v.cpu().numpy() for k, v in d.items()}
        out["leaf_index"] = out["leaf_index"].view(np.uint32)
        out["codes"] = out["codes"].view(np.uint64)
        return out

    def spatial_crs(self, preds, kind=PRED_SPHERE, sort_predicates=True, buffer_size=0):
        stride = {0: 4, 1: 6, 2: 3, 3: 6}[kind]
        p = abx.intersects(_dev(np.asarray(preds, np.float32).reshape(-1, stride)), kind)
        try:
            idx, off = self.bvh.query(self.space, p, abx.TraversalPolicy(buffer_size, sort_predicates))
        except abx.SearchException as e:
            raise SearchException(str(e))
        self.space.fence()
        return off.cpu().numpy(), idx.cpu().numpy().view(np.uint32)

    def spatial_count(self, preds, kind=PRED_SPHERE, limit=0):
        stride = {0: 4, 1: 6, 2: 3, 3: 6}[kind]
        p = abx.intersects(_dev(np.asarray(preds, np.float32).reshape(-1, stride)), kind)
        c = self.bvh.count(self.space, p, limit)
        self.space.fence()
        return c.cpu().numpy()

    def nearest_crs(self, pts, k, sort_predicates=True):
        kk = torch.as_tensor(np.asarray(k, np.int32)).cuda() if np.ndim(k) > 0 else int(k)
        p = abx.nearest(_dev(np.asarray(pts, np.float32).reshape(-1, 3)), kk)
        idx, off, dist = self.bvh.query(self.space, p, abx.TraversalPolicy(0, sort_predicates), return_distances=True)
        self.space.fence()
        return off.cpu().numpy(), idx.cpu().numpy().view(np.uint32), dist.cpu().numpy()

    def nearest_geom_crs(self, preds, kind, k, sort_predicates=True):
        stride = {0: 4, 1: 6, 2: 3, 3: 6}[kind]
        p = abx.nearest(_dev(np.asarray(preds, np.float32).reshape(-1, stride)), int(k), kind)
        idx, off, dist = self.bvh.query(self.space, p, abx.TraversalPolicy(0, sort_predicates), return_distances=True)
        self.space.fence()
        return off.cpu().numpy(), idx.cpu().numpy().view(np.uint32), dist.cpu().numpy()

    def half_pairs(self, r):
        pairs = self.bvh.half_traversal_pairs(self.space, r)
        self.space.fence()
        return pairs.cpu().numpy().view(np.uint32)


class CudaEngine:
    name = "cuda"

    def __init__(self):
        self.space = abx.ExecutionSpace()

    def ensure(self):
        pass

    def build(self, prims, kind=PRIM_POINT):
        stride = {0: 3, 1: 6, 2: 9}[kind]
        v = _dev(np.asarray(prims, np.float32).reshape(-1, stride))
        bvh = abx.BoundingVolumeHierarchy(self.space, v, kind)
        self.space.fence()
        return _CudaTree(self.space, bvh, v.shape[0], kind)

    def from_sorted_codes(self, prims, codes, kind=PRIM_POINT):
        stride = {0: 3, 1: 6, 2: 9}[kind]
        v = _dev(np.asarray(prims, np.float32).reshape(-1, stride))
        c = torch.as_tensor(np.asarray(codes, np.uint64).view(np.int64)).cuda()
        bvh = abx.BoundingVolumeHierarchy._from_sorted_codes(self.space, v, c, kind)
        return _CudaTree(self.space, bvh, v.shape[0], kind)

    def scene_bounds(self, prims, kind=PRIM_POINT):
        stride = {0: 3, 1: 6, 2: 9}[kind]
        v = _dev(np.asarray(prims, np.float32).reshape(-1, stride))
        out = torch.empty(6, dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib().abx_scene_bounds(self.space.handle, kind, C.c_void_p(v.data_ptr()), v.shape[0],
                                               C.c_void_p(out.data_ptr())))
        self.space.fence()
        return out.cpu().numpy()

    def morton64_codes(self, prims, bounds6, kind=PRIM_POINT):
        stride = {0: 3, 1: 6, 2: 9}[kind]
        v = _dev(np.asarray(prims, np.float32).reshape(-1, stride))
        b = _dev(np.asarray(bounds6, np.float32))
        out = torch.empty(v.shape[0], dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib().abx_morton64(self.space.handle, kind, C.c_void_p(v.data_ptr()), v.shape[0],
                                           C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr())))
        self.space.fence()
        return out.cpu().numpy().view(np.uint64)

    def sort_u64(self, keys):
        k = torch.as_tensor(np.asarray(keys, np.uint64).view(np.int64)).cuda().clone()
        perm = torch.empty(k.shape[0], dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().abx_sort_u64(self.space.handle, C.c_void_p(k.data_ptr()), C.c_void_p(perm.data_ptr()),
                                           k.shape[0]))
        self.space.fence()
        return k.cpu().numpy().view(np.uint64), perm.cpu().numpy().view(np.uint32)

    def sort_u32(self, keys):
        k = torch.as_tensor(np.asarray(keys, np.uint32).view(np.int32)).cuda().clone()
        perm = torch.empty(k.shape[0], dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().abx_sort_u32(self.space.handle, C.c_void_p(k.data_ptr()), C.c_void_p(perm.data_ptr()),
                                           k.shape[0]))
        self.space.fence()
        return k.cpu().numpy().view(np.uint32), perm.cpu().numpy().view(np.uint32)

    def dbscan(self, xyz, eps, minpts, impl=0, algo=0):
        x = _dev(np.asarray(xyz, np.float32).reshape(-1, 3))
        try:
            labels = abx.dbscan(self.space, x, eps, minpts, abx.DBSCANParameters(impl, algo))
        except abx.SearchException as e:
            raise SearchException(str(e))
        self.space.fence()
        return labels.cpu().numpy()
