"""ArborX::BruteForce (spatial/ArborX_BruteForce.hpp, detail/ArborX_BruteForceImpl.hpp:40-233) against the oracle's
tree: identical result sets for spatial predicates, identical distance rows for nearest predicates
(test/tstQueryTreeComparisonWithBoost-style: the exhaustive search IS the independent implementation)."""
import numpy as np
import pytest
import torch

from tests import clouds
from tests.engines import rows_of

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.mark.parametrize("n", [0, 1, 5, 700, 5000])
def test_brute_force_matches_oracle(n):
    import arborx_b200 as abx
    import oracle
    space = abx.ExecutionSpace()
    pts = clouds.filled_box(101, max(n, 8))[:n]
    boxes = np.concatenate([pts, pts + clouds.uniform01(102, max(n, 8))[:n] * F(0.8)], 1).astype(F)
    q = clouds.filled_box(103, 900)
    spheres = np.concatenate([q, np.full((900, 1), 2.0, F)], 1).astype(F)
    qboxes = np.concatenate([q, q + F(1.7)], 1).astype(F)
    rays = clouds.ball_rays(104, 900) * np.array([4, 4, 4, 1, 1, 1], F)
    for prims, kind in ((pts, abx.POINT), (boxes, abx.BOX)):
        bf = abx.BruteForce(space, torch.from_numpy(prims).cuda().reshape(-1, 3 if kind == abx.POINT else 6), kind)
        assert bf.size() == n and bf.empty() == (n == 0)
        ref = oracle.Tree(prims, kind)
        if n:
            assert np.array_equal(bf.bounds().numpy(), ref.bounds())
        cases = [(spheres, abx.SPHERE_PRED), (qboxes, abx.BOX_PRED), (q, abx.POINT_PRED)]
        if kind == abx.BOX:
            cases.append((rays, abx.RAY_PRED))
        for preds, pk in cases:
            idx, off = bf.query(space, abx.intersects(torch.from_numpy(preds).cuda(), pk))
            roff, ridx = ref.spatial_crs(preds, pk)
            assert np.array_equal(off.cpu().numpy(), roff)
            assert rows_of(off.cpu().numpy(), idx.cpu().numpy()) == rows_of(roff, ridx)
        for k in (1, 6, 40):
            idx, off, d = bf.query(space, abx.nearest(torch.from_numpy(q).cuda(), k), return_distances=True)
            roff, ridx, rd = ref.nearest_crs(q, k)
            assert np.array_equal(off.cpu().numpy(), roff)
            assert np.array_equal(d.cpu().numpy(), rd)


def test_brute_force_rejects_triangles():
    import arborx_b200 as abx
    space = abx.ExecutionSpace()
    with pytest.raises((ValueError, KeyError)):
        abx.BruteForce(space, torch.zeros((3, 9), device="cuda"), abx.TRIANGLE)
