"""Config 5 (triangulated surface): the icosphere generator restated from the reference, trees over its triangles in
both input forms (flat corners, vertex-index triples), nearest(point, 1) and intersects(ray) against the oracle."""
import numpy as np
import pytest
import torch

from tests import clouds

F = np.float32


def test_icosphere_generator():
    """generator.hpp:29-71,176-321: 20 * 4^r triangles, 10 * 4^r + 2 vertices on the unit sphere, a closed surface
    (every edge shared by exactly two triangles), consistent orientation-free connectivity."""
    for r in (0, 1, 3, 5):
        v, t = clouds.icosphere(r)
        assert t.shape == (20 * 4 ** r, 3) and v.shape == (10 * 4 ** r + 2, 3)
        assert np.allclose(np.linalg.norm(v.astype(np.float64), axis=1), 1.0, atol=1e-6)
        e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), 1)
        _, c = np.unique(e, axis=0, return_counts=True)
        assert (c == 2).all()
        assert (t[:, 0] != t[:, 1]).all() and (t[:, 1] != t[:, 2]).all() and (t[:, 0] != t[:, 2]).all()
    # radius scaling
    v2, _ = clouds.icosphere(2, radius=3.0)
    assert np.allclose(np.linalg.norm(v2.astype(np.float64), axis=1), 3.0, atol=1e-5)


def test_torch_cloud_generator_matches_numpy():
    a = clouds.filled_box(7, 100_000)
    b = clouds.filled_box_torch(7, 100_000, "cpu").numpy()
    assert np.array_equal(a, b)
    c = clouds.filled_box_torch(7, 1000, "cpu", a=float(F(np.cbrt(100000.0))), first=500).numpy()
    assert np.array_equal(a[500:1500], c)


@pytest.mark.gpu
@pytest.mark.parametrize("refinements", [0, 4])
def test_indexed_triangles_equal_flat_triangles(refinements):
    import arborx_b200 as abx
    import oracle
    space = abx.ExecutionSpace()
    v, t = clouds.icosphere(refinements)
    soup = clouds.triangle_soup(v, t)
    flat = abx.BoundingVolumeHierarchy(space, torch.from_numpy(soup).cuda(), abx.TRIANGLE)
    indexed = abx.BoundingVolumeHierarchy.from_indexed_triangles(space, torch.from_numpy(v).cuda(), torch.from_numpy(t).cuda())
    assert indexed.size() == flat.size() == len(t)
    assert torch.equal(indexed.bounds(), flat.bounds())
    a, b = flat.export_reference_layout(space), indexed.export_reference_layout(space)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    ref = oracle.Tree(soup, oracle.PRIM_TRI)
    q = np.concatenate([clouds.shell_points(81, 2000), clouds.filled_box(82, 500) * F(0.2)]).astype(F)
    idx, off, d = indexed.query(space, abx.nearest(torch.from_numpy(q).cuda(), 1), return_distances=True)
    roff, ridx, rd = ref.nearest_crs(q, 1)
    assert np.array_equal(off.cpu().numpy(), roff) and np.array_equal(d.cpu().numpy(), rd)
    rays = clouds.ball_rays(83, 2000)
    idx, off = indexed.query(space, abx.intersects(torch.from_numpy(rays).cuda(), abx.RAY_PRED))
    roff, ridx = ref.spatial_crs(rays, oracle.PRED_RAY)
    assert np.array_equal(off.cpu().numpy(), roff)
    row = np.repeat(np.arange(len(rays)), np.diff(roff))
    gi = idx.cpu().numpy().view(np.uint32)
    assert np.array_equal(gi[np.lexsort((gi, row))], ridx[np.lexsort((ridx, row))])
    with pytest.raises(ValueError):
        bad = t.copy()
        bad[0, 0] = len(v)
        abx.BoundingVolumeHierarchy.from_indexed_triangles(space, torch.from_numpy(v).cuda(), torch.from_numpy(bad).cuda())


@pytest.mark.gpu
def test_ray_triangle_tolerance_hits_fine_mesh():
    """On a fine mesh the reference's ray - triangle test (ArborX_Ray.hpp:266-417) accepts hits within its 1e-7
    tolerance just outside a triangle, i.e. on triangles whose own bounding box the ray may miss.  The reference hands
    a leaf's value to the predicate whenever the leaf's parent is visited (TreeTraversal.hpp:97-119), so those hits
    are part of its answer: the row sets must equal the oracle's, and some rays must indeed report more triangles
    than a leaf-box pre-test would allow (which is what the 20M-triangle bench workload exposed)."""
    import arborx_b200 as abx
    import oracle
    space = abx.ExecutionSpace()
    v, t = clouds.icosphere(8)  # 1.3M triangles
    soup = clouds.triangle_soup(v, t)
    bvh = abx.BoundingVolumeHierarchy(space, torch.from_numpy(soup).cuda(), abx.TRIANGLE)
    ref = oracle.Tree(soup, oracle.PRIM_TRI)
    rays = clouds.ball_rays(0x5EED0053, 150_000)
    idx, off = bvh.query(space, abx.intersects(torch.from_numpy(rays).cuda(), abx.RAY_PRED))
    roff, ridx = ref.spatial_crs(rays, oracle.PRED_RAY)
    assert np.array_equal(off.cpu().numpy(), roff)
    row = np.repeat(np.arange(len(rays)), np.diff(roff))
    gi = idx.cpu().numpy().view(np.uint32)
    assert np.array_equal(gi[np.lexsort((gi, row))], ridx[np.lexsort((ridx, row))])
    # the tolerance hits exist in this sample: a hit triangle whose box the ray misses
    lo = soup.reshape(-1, 3, 3).min(1)
    hi = soup.reshape(-1, 3, 3).max(1)
    o, d = rays[row, :3].astype(np.float64), rays[row, 3:].astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (lo[ridx] - o) / d
        t1 = (hi[ridx] - o) / d
    tn = np.nanmax(np.minimum(t0, t1), axis=1)
    tf = np.nanmin(np.maximum(t0, t1), axis=1)
    assert int((tn > tf).sum()) > 0
    # sphere predicates over the same tree
    c = clouds.shell_points(91, 20_000)
    spheres = np.concatenate([c, np.full((len(c), 1), 0.01, F)], 1).astype(F)
    idx, off = bvh.query(space, abx.intersects(torch.from_numpy(spheres).cuda()))
    roff, ridx = ref.spatial_crs(spheres, oracle.PRED_SPHERE)
    assert np.array_equal(off.cpu().numpy(), roff)
