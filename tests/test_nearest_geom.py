"""nearest(Box | Sphere | Ray, k) (spatial/detail/ArborX_Predicates.hpp:58-80 with the distances of
geometry/algorithms/ArborX_Distance.hpp:83-108,166-209 and geometry/ArborX_Ray.hpp:433-444): both engines against
float32 brute force with the reference's operation order, the CUDA engine against the oracle for rays."""
import numpy as np
import pytest

from tests import brute, clouds
from tests.engines import PRED_BOX, PRED_POINT, PRED_RAY, PRED_SPHERE, PRIM_BOX, PRIM_POINT

F = np.float32


def _box_box_dist(qb, boxes):
    """distance(Box a = query, Box b) (ArborX_Distance.hpp:166-197), float32, axis by axis."""
    a_lo, a_hi = qb[:, None, 0:3], qb[:, None, 3:6]
    b_lo, b_hi = boxes[None, :, 0:3], boxes[None, :, 3:6]
    delta = np.where(a_lo > b_hi, a_lo - b_hi, np.where(b_lo > a_hi, b_lo - a_hi, F(0))).astype(F)
    d2 = np.zeros(delta.shape[:2], F)
    for d in range(3):
        d2 = (d2 + (delta[..., d] * delta[..., d]).astype(F)).astype(F)
    return np.sqrt(d2).astype(F)


@pytest.mark.parametrize("n", [1, 2, 7, 3000])
@pytest.mark.parametrize("k", [1, 4, 20])
def test_nearest_box_and_sphere_vs_bruteforce(engine, n, k):
    pts = clouds.filled_box(61, max(n, 8))[:n]
    boxes = np.concatenate([pts, pts + clouds.uniform01(62, max(n, 8))[:n] * F(0.7)], 1).astype(F)
    q = clouds.filled_box(63, 300) * F(1.1)
    spheres = np.concatenate([q, clouds.uniform01(64, 300, 1) * F(2.0)], 1).astype(F)
    qboxes = np.concatenate([q, q + clouds.uniform01(65, 300) * F(1.5)], 1).astype(F)
    row = min(k, n)
    for prims, kind in ((pts, PRIM_POINT), (boxes, PRIM_BOX)):
        tree = engine.build(prims, kind)
        as_boxes = boxes if kind == PRIM_BOX else np.concatenate([pts, pts], 1)
        # sphere: max(distance(centre, X) - r, 0)
        off, idx, d = tree.nearest_geom_crs(spheres, PRED_SPHERE, k)
        assert np.array_equal(off, np.arange(301) * row)
        dc = brute.dist_point_box(q, as_boxes) if kind == PRIM_BOX else brute.dist_point_point(q, pts)
        ds = np.maximum((dc - spheres[:, 3:4]).astype(F), F(0))
        assert np.array_equal(d.reshape(300, row), np.sort(ds, 1)[:, :row])
        assert np.array_equal(np.take_along_axis(ds, idx.reshape(300, row).astype(np.int64), 1), d.reshape(300, row))
        # box
        off, idx, d = tree.nearest_geom_crs(qboxes, PRED_BOX, k)
        db = _box_box_dist(qboxes, as_boxes)
        assert np.array_equal(d.reshape(300, row), np.sort(db, 1)[:, :row])
        assert np.array_equal(np.take_along_axis(db, idx.reshape(300, row).astype(np.int64), 1), d.reshape(300, row))
        # point through the same entry point
        off, idx, d = tree.nearest_geom_crs(q, PRED_POINT, k)
        assert np.array_equal(d.reshape(300, row), np.sort(dc, 1)[:, :row])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 50, 20_000])
@pytest.mark.parametrize("k", [1, 3, 10])
def test_nearest_ray_cuda_vs_oracle(n, k):
    """Ray casting with nearest queries: the k boxes a ray enters first (rows short when it hits fewer)."""
    from tests.engines import CudaEngineLazy, OracleEngine
    cuda, orc = CudaEngineLazy(), OracleEngine()
    lo = clouds.filled_box(71, max(n, 8))[:n]
    boxes = np.concatenate([lo, lo + clouds.uniform01(72, max(n, 8))[:n] * F(1.5)], 1).astype(F)
    a = F(np.cbrt(float(max(n, 8))))
    rays = clouds.ball_rays(73, 2000) * np.array([a, a, a, 1, 1, 1], F)
    rays[::11, 3:] = np.array([1, 0, 0], F)  # axis-aligned directions (zero components)
    rays[::13, 3:] = np.array([0, 0, -2], F)
    for prims, kind in ((boxes, PRIM_BOX), (lo, PRIM_POINT)):
        go, gi, gd = cuda.build(prims, kind).nearest_geom_crs(rays, PRED_RAY, k)
        ro, ri, rd = orc.build(prims, kind).nearest_geom_crs(rays, PRED_RAY, k)
        assert np.array_equal(go, ro)
        assert np.array_equal(gd, rd)
        # indices may differ only among equal distances: the reported distance must be the box's own
        same = gi == ri
        if not same.all():
            row = np.repeat(np.arange(len(rays)), np.diff(ro))
            for j in np.nonzero(~same)[0]:
                grp = (row == row[j]) & (rd == rd[j])
                assert sorted(gi[grp]) == sorted(ri[grp])
    if n >= 50:
        assert int(np.diff(ro).max()) > 0 or kind == PRIM_POINT
