"""The two node walks of the spatial kernels.  User-facing trees carry 4-wide quantised records (Wide64, written
at build time) that the spatial kernels walk by default; the exact Node64 walk remains for small trees and -- chosen
on the device -- for trees whose boxes cannot be quantised (non-finite coordinates).  Both must give the oracle's
result sets; the quantised boxes only cull, every reported leaf is tested exactly."""
import numpy as np
import pytest

from tests import brute, clouds
from tests.engines import PRED_BOX, PRED_SPHERE, PRIM_BOX, PRIM_POINT, CudaEngineLazy, OracleEngine, rows_of

pytestmark = pytest.mark.gpu
F = np.float32


def _same_rows(a, b):
    assert np.array_equal(a[0], b[0])
    assert rows_of(*a) == rows_of(*b)


@pytest.mark.parametrize("n", [65, 300, 5000, 200_000])
def test_wide_walk_matches_oracle(n):
    cuda, orc = CudaEngineLazy(), OracleEngine()
    pts = clouds.filled_box(41, n)
    q = clouds.filled_box(42, 3000)
    r = clouds.bvh_driver_radius(10)
    sp = np.concatenate([q, np.full((len(q), 1), r, F)], 1).astype(F)
    _same_rows(cuda.build(pts).spatial_crs(sp), orc.build(pts).spatial_crs(sp))
    # boxes as primitives and as predicates; degenerate extents (scale 0 on an axis)
    boxes = np.concatenate([pts, pts + clouds.uniform01(43, n) * F(0.5)], 1).astype(F)
    boxes[::7, 3] = boxes[::7, 0]
    qb = np.concatenate([q, q + F(1.5)], 1).astype(F)
    _same_rows(cuda.build(boxes, PRIM_BOX).spatial_crs(qb, PRED_BOX), orc.build(boxes, PRIM_BOX).spatial_crs(qb, PRED_BOX))
    flat = pts.copy()
    flat[:, 2] = F(3.25)  # a planar cloud: every node has zero extent in z
    _same_rows(cuda.build(flat).spatial_crs(sp), orc.build(flat).spatial_crs(sp))


def test_unquantisable_tree_falls_back_on_device():
    """Boxes with infinite extents cannot be quantised: the converter flags the tree and the kernels take the
    Node64 walk, without the host ever reading the flag."""
    cuda = CudaEngineLazy()
    n = 4000
    lo = clouds.filled_box(51, n)
    boxes = np.concatenate([lo, lo + F(0.3)], 1).astype(F)
    boxes[10, 3:] = np.inf
    boxes[11, :3] = -np.inf
    tree = cuda.build(boxes, PRIM_BOX)
    q = clouds.filled_box(52, 500)
    sp = np.concatenate([q, np.full((len(q), 1), 1.0, F)], 1).astype(F)
    off, idx = tree.spatial_crs(sp, PRED_SPHERE)
    # brute force with the reference's float32 operation order
    with np.errstate(invalid="ignore"):
        mask = brute.spheres_vs_boxes(sp, boxes)
    expect = [sorted(np.nonzero(m)[0].tolist()) for m in mask]
    assert rows_of(off, idx) == expect
    assert sum(10 in row for row in expect) > 0  # the half-infinite box is found
