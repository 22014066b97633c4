"""Two interchangeable back ends for the golden/known-answer tests: the CPU
oracle (checker) and the CUDA product called through its C ABI (via the
arborx_b200 Python mirror).  Both expose numpy-in / numpy-out methods."""
import numpy as np

import oracle

PRIM_POINT, PRIM_BOX, PRIM_TRI = 0, 1, 2
PRED_SPHERE, PRED_BOX, PRED_POINT, PRED_RAY = 0, 1, 2, 3


class SearchException(Exception):
    pass


class OracleEngine:
    name = "oracle"

    def ensure(self):
        oracle.lib()

    def build(self, prims, kind=PRIM_POINT):
        return _OracleTree(prims, kind)

    def from_sorted_codes(self, prims, codes, kind=PRIM_POINT):
        return _OracleTree(prims, kind, codes)

    def scene_bounds(self, prims, kind=PRIM_POINT):
        return oracle.scene_bounds(prims, kind)

    def morton64_codes(self, prims, bounds6, kind=PRIM_POINT):
        return oracle.morton64_codes(prims, bounds6, kind)

    def sort_u64(self, keys):
        return oracle.sort_u64(keys)

    def dbscan(self, xyz, eps, minpts, impl=0, algo=0):
        try:
            return oracle.dbscan(xyz, eps, minpts, impl, algo)
        except ValueError as e:
            raise SearchException(str(e))


class _OracleTree(oracle.Tree):
    def __init__(self, prims, kind, codes=None):
        super().__init__(prims, kind, codes)

    def spatial_crs(self, preds, kind=PRED_SPHERE, sort_predicates=True, buffer_size=0):
        try:
            return super().spatial_crs(preds, kind, sort_predicates, buffer_size)
        except RuntimeError as e:
            raise SearchException(str(e))


class CudaEngineLazy:
    """Resolved on first use so that CPU-only collection never imports torch.cuda."""
    name = "cuda"
    _impl = None

    def ensure(self):
        if CudaEngineLazy._impl is None:
            from tests.cuda_engine import CudaEngine
            CudaEngineLazy._impl = CudaEngine()

    def __getattr__(self, item):
        self.ensure()
        return getattr(CudaEngineLazy._impl, item)


def rows_of(offsets, indices):
    offsets = np.asarray(offsets)
    return [sorted(int(x) for x in indices[offsets[i]:offsets[i + 1]]) for i in range(len(offsets) - 1)]
