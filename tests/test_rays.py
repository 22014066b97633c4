"""Ray predicates (geometry/ArborX_Ray.hpp): the reference's known answers (test/tstRay.cpp, extracted to
tests/golden/ray_vectors.json by tests/golden/make_ray_vectors.py) against the oracle (CPU) and through
one- and two-leaf trees on both engines; random rays against the oracle on the GPU."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests import clouds
from tests.engines import PRED_RAY, PRIM_BOX, PRIM_TRI, rows_of

F = np.float32
VEC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ray_vectors.json")))


def test_ray_known_answers_oracle():
    import oracle
    L = oracle.lib()

    def fp(a):
        a = np.ascontiguousarray(a, F)
        return a, a.ctypes.data_as(C.POINTER(C.c_float))

    assert len(VEC) == 139
    for v in VEC:
        r, rp = fp(v["ray"])
        g, gp = fp(v["geom"])
        assert bool(L.orc_intersects(PRED_RAY, rp, PRIM_BOX if v["kind"] == "box" else PRIM_TRI, gp)) == v["hit"], v


def test_ray_known_answers_through_trees(engine):
    for kind, prim in (("box", PRIM_BOX), ("triangle", PRIM_TRI)):
        groups = {}
        for v in VEC:
            if v["kind"] == kind:
                groups.setdefault(tuple(v["geom"]), []).append(v)
        for geom, vs in groups.items():
            rays = np.array([v["ray"] for v in vs], F)
            expect = [[0] if v["hit"] else [] for v in vs]
            # single-leaf tree and a two-leaf tree with an unreachable far-away primitive
            far = np.array(geom, F) + F(1000.0)
            for prims in (np.array([geom], F), np.array([geom, far], F)):
                t = engine.build(prims, prim)
                off, idx = t.spatial_crs(rays, PRED_RAY)
                got = rows_of(off, idx)
                assert [[i for i in r if i == 0] for r in got] == expect, (kind, geom)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["boxes", "triangles"])
def test_random_rays_vs_oracle(kind):
    from tests.engines import CudaEngineLazy, OracleEngine
    cuda, orc = CudaEngineLazy(), OracleEngine()
    cuda.ensure()
    n, q = 30_000, 5000
    if kind == "boxes":
        lo = clouds.uniform01(3, n)
        prims = np.concatenate([lo, lo + clouds.uniform01(4, n) * F(0.02)], 1).astype(F)
        pk = PRIM_BOX
    else:
        a = clouds.uniform01(5, n)
        prims = np.concatenate([a, a + (clouds.uniform01(6, n) - F(0.5)) * F(0.05),
                                a + (clouds.uniform01(7, n) - F(0.5)) * F(0.05)], 1).astype(F)
        pk = PRIM_TRI
    origins = clouds.uniform01(8, q)
    dirs = clouds.uniform01(9, q) - F(0.5)
    dirs[::7, 0] = 0  # axis-aligned components exercise the +-inf branches
    dirs[::11, 1:] = 0
    dirs[np.all(dirs == 0, 1)] = 1
    rays = np.concatenate([origins, dirs], 1).astype(F)
    tc, to = cuda.build(prims, pk), orc.build(prims, pk)
    oc, ic = tc.spatial_crs(rays, PRED_RAY)
    oo, io = to.spatial_crs(rays, PRED_RAY)
    assert np.array_equal(oc, oo)
    assert rows_of(oc, ic) == rows_of(oo, io)
    assert oo[-1] > q  # the test is not vacuous
