"""Brute-force float32 checkers with the reference's operation order
(geometry/algorithms/ArborX_Distance.hpp:54-70): independent of any tree.
Stands in for the Boost R-tree comparisons of
test/tstQueryTreeComparisonWithBoost.cpp (Boost is not in this image)."""
import numpy as np

F = np.float32


def dist_point_point(a, b):
    """a: [m,3] queries, b: [n,3] points -> [m,n] float32 distances."""
    a = a.astype(F)[:, None, :]
    b = b.astype(F)[None, :, :]
    t = b - a
    d2 = (t[..., 0] * t[..., 0]).astype(F)
    d2 = (d2 + (t[..., 1] * t[..., 1]).astype(F)).astype(F)
    d2 = (d2 + (t[..., 2] * t[..., 2]).astype(F)).astype(F)
    return np.sqrt(d2).astype(F)


def dist_point_box(p, boxes):
    p = p.astype(F)[:, None, :]
    lo = boxes[None, :, 0:3].astype(F)
    hi = boxes[None, :, 3:6].astype(F)
    c = np.where(p < lo, lo, np.where(p > hi, hi, p)).astype(F)
    t = c - p
    d2 = (t[..., 0] * t[..., 0]).astype(F)
    d2 = (d2 + (t[..., 1] * t[..., 1]).astype(F)).astype(F)
    d2 = (d2 + (t[..., 2] * t[..., 2]).astype(F)).astype(F)
    return np.sqrt(d2).astype(F)


def spheres_vs_points(spheres, pts):
    d = dist_point_point(spheres[:, :3], pts)
    return d <= spheres[:, 3:4].astype(F)


def spheres_vs_boxes(spheres, boxes):
    return dist_point_box(spheres[:, :3], boxes) <= spheres[:, 3:4].astype(F)


def boxes_vs_boxes(qb, boxes):
    q_lo, q_hi = qb[:, None, 0:3], qb[:, None, 3:6]
    lo, hi = boxes[None, :, 0:3], boxes[None, :, 3:6]
    return ~np.any((q_lo > hi) | (q_hi < lo), axis=2)


def rows_from_mask(mask):
    return [sorted(np.nonzero(r)[0].tolist()) for r in mask]


def knn_check(dmat, k, offsets, indices, dists, rtol=1e-6):
    """Tie-tolerant kNN check: row sizes, ascending order, reported distances
    match the brute-force distance of the reported index, and the multiset of
    distances equals the k smallest brute-force distances."""
    q, n = dmat.shape
    for i in range(q):
        row = indices[offsets[i]:offsets[i + 1]]
        rd = dists[offsets[i]:offsets[i + 1]]
        kk = min(k if np.ndim(k) == 0 else int(k[i]), n)
        assert len(row) == max(kk, 0), (i, len(row), kk)
        assert len(set(row.tolist())) == len(row)
        assert np.all(np.diff(rd) >= 0)
        true = np.sort(dmat[i])[:kk]
        assert np.allclose(rd, true, rtol=rtol, atol=0), (i, rd, true)
        assert np.allclose(dmat[i, row], rd, rtol=rtol, atol=0)
