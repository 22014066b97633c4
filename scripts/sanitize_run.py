"""Small run through every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_run.py
    compute-sanitizer --tool racecheck python scripts/sanitize_run.py 6000
Sizes are small (sanitizer slows kernels 10-100x); results are cross-checked against the oracle where cheap."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
import oracle  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000
F = np.float32
space = abx.ExecutionSpace()
pts = clouds.filled_box(1, n)
qp = clouds.filled_box(2, n // 2)
r = float(clouds.bvh_driver_radius(10))
spheres = np.concatenate([qp, np.full((len(qp), 1), r, F)], 1)
boxes_q = np.concatenate([qp - F(1.0), qp + F(1.0)], 1)
prim_boxes = np.concatenate([pts, pts + F(0.7)], 1)
tris = np.concatenate([pts, pts + F(0.5) * clouds.uniform01(3, n), pts + F(0.5) * clouds.uniform01(4, n)], 1)

for name, prims, kind in (("points", pts, 0), ("boxes", prim_boxes, 1), ("triangles", tris, 2)):
    bvh = abx.BoundingVolumeHierarchy(space, torch.from_numpy(prims).cuda(), kind)
    idx, off = bvh.query(space, abx.intersects(torch.from_numpy(spheres).cuda()))
    ref = oracle.Tree(prims, kind)
    roff, ridx = ref.spatial_crs(spheres, 0)
    assert np.array_equal(off.cpu().numpy(), roff), name
    if kind != 2:
        idx2, off2 = bvh.query(space, abx.intersects(torch.from_numpy(boxes_q).cuda(), 1))
        roff2, _ = ref.spatial_crs(boxes_q, 1)
        assert np.array_equal(off2.cpu().numpy(), roff2), name
    for k in (1, 10, 24, 40):
        kidx, koff, kd = bvh.query(space, abx.nearest(torch.from_numpy(qp).cuda(), k), return_distances=True)
        _, _, rd = ref.nearest_crs(qp, k)
        assert np.allclose(kd.cpu().numpy(), rd, rtol=1e-6), (name, k)
    cnt = bvh.count(space, abx.intersects(torch.from_numpy(spheres).cuda()), limit=5)
    assert int(cnt.max()) <= 5
    print("ok", name, int(off[-1]))

# sort paths: fix-up, escalation, plain
from tests.cuda_engine import CudaEngine  # noqa: E402
eng = CudaEngine()
rng = np.random.default_rng(0)
for m in (5000, n):
    keys = rng.integers(0, 2 ** 63, m, dtype=np.uint64)
    k, p = eng.sort_u64(keys)
    assert np.array_equal(p, np.argsort(keys, kind="stable").astype(np.uint32))
    keys[: m // 2] = (keys[0] >> np.uint64(39)) << np.uint64(39) | rng.integers(0, 2 ** 39, m // 2, dtype=np.uint64)
    k, p = eng.sort_u64(keys)
    assert np.array_equal(p, np.argsort(keys, kind="stable").astype(np.uint32))
print("ok sort")

# DBSCAN, both implementations
cl = clouds.clustered(5, n, domain=1000.0, spread=5.0)
for impl in (0, 1):
    for minpts in (2, 5):
        lab = abx.dbscan(space, torch.from_numpy(cl).cuda(), 6.0, minpts, abx.DBSCANParameters(impl, 0)).cpu().numpy()
        assert oracle.dbscan_verify(cl, 6.0, minpts, lab, 0) == 0, (impl, minpts)
print("ok dbscan")

# DistributedTree building blocks on one rank
import torch.distributed as dist  # noqa: E402
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29577")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
from arborx_b200.distributed import DistributedTree  # noqa: E402
from arborx_b200.distributed_dbscan import dbscan as ddbscan  # noqa: E402
tree = DistributedTree(dist.group.WORLD, space, torch.from_numpy(pts).cuda())
v, o = tree.query(space, abx.intersects(torch.from_numpy(spheres).cuda()))
kv, ko, kd = tree.query(space, abx.nearest(torch.from_numpy(qp).cuda(), 7), return_distances=True)
lab = ddbscan(dist.group.WORLD, space, torch.from_numpy(cl).cuda(), 6.0, 5)
torch.cuda.synchronize()
dist.destroy_process_group()
print("ok distributed", int(o[-1]), int(ko[-1]))
# round-2 kernel families: rays over triangles (leaves tested without a box), nearest(geometry), BruteForce, MST /
# dendrogram, and the multi-rank exchange with in-process ranks
rays = clouds.ball_rays(7, 4000)
v_ico, t_ico = clouds.icosphere(4)
soup = clouds.triangle_soup(v_ico, t_ico)
tb = abx.BoundingVolumeHierarchy(space, torch.from_numpy(soup).cuda(), abx.TRIANGLE)
ridx, roff = tb.query(space, abx.intersects(torch.from_numpy(rays).cuda(), abx.RAY_PRED))
o_off, _ = oracle.Tree(soup, oracle.PRIM_TRI).spatial_crs(rays, oracle.PRED_RAY)
assert np.array_equal(roff.cpu().numpy(), o_off)
bb = abx.BoundingVolumeHierarchy(space, torch.from_numpy(prim_boxes).cuda(), 1)
gq = np.concatenate([qp[:2000] - F(0.5), qp[:2000] + F(0.5)], 1)
gi, go, gd = bb.query(space, abx.nearest(torch.from_numpy(gq).cuda(), 5, kind=abx.BOX_PRED),
                      return_distances=True)
brute = abx.BruteForce(space, torch.from_numpy(pts[:5000]).cuda())
bi, bo = brute.query(space, abx.intersects(torch.from_numpy(spheres[:3000]).cuda()))
ki, ko2 = brute.query(space, abx.nearest(torch.from_numpy(qp[:3000]).cuda(), 6))
print("ok rays / nearest(geometry) / brute force", int(roff[-1]), int(go[-1]), int(bo[-1]))
for k in (1, 4):
    mst = abx.MinimumSpanningTree(space, torch.from_numpy(cl).cuda(), k)
    e, w = oracle.mst(cl, k)
    assert np.array_equal(np.sort(mst.weights.cpu().numpy()), np.sort(w)), k
d = abx.hdbscan(space, torch.from_numpy(cl[:5000]).cuda(), 3)
torch.cuda.synchronize()
print("ok mst / hdbscan", mst.iterations)

import threading  # noqa: E402
from arborx_b200.distributed import Communicator  # noqa: E402
from arborx_b200.distributed_dbscan import dbscan as ddbscan2  # noqa: E402
world = 3
comms = Communicator.local_group(world)
errs = [None] * world


def worker(rk):
    try:
        torch.cuda.set_device(0)
        sp = abx.ExecutionSpace(torch.cuda.Stream())
        with torch.cuda.stream(sp.stream):
            mine = torch.from_numpy(np.ascontiguousarray(pts[rk::world])).cuda()
            tr = DistributedTree(comms[rk], sp, mine)
            tr.query(sp, abx.intersects(torch.from_numpy(np.ascontiguousarray(spheres[rk::world])).cuda()))
            tr.query(sp, abx.nearest(torch.from_numpy(np.ascontiguousarray(qp[rk::world])).cuda(), 7), return_distances=True)
            tr.query(sp, abx.nearest(torch.from_numpy(np.ascontiguousarray(qp[rk::world])), 7))  # host form
            ddbscan2(comms[rk], sp, torch.from_numpy(np.ascontiguousarray(cl[rk::world])).cuda(), 6.0, 5)
            sp.fence()
    except BaseException:
        import traceback
        errs[rk] = traceback.format_exc()


ths = [threading.Thread(target=worker, args=(rk,), daemon=True) for rk in range(world)]
[t.start() for t in ths]
[t.join(timeout=600) for t in ths]
assert not any(t.is_alive() for t in ths) and not any(errs), errs
print("ok in-process ranks")
print("SANITIZE RUN OK")
