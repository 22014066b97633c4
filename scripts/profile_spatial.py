"""One build + two radius queries (and optionally kNN) at n points for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'spatialKernel|nearestKernel' -s 2 -c 2 \
        -o gpurun_out/prof_spatial python scripts/profile_spatial.py [n] [knn]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
knn = len(sys.argv) > 2 and sys.argv[2] == "knn"
space = abx.ExecutionSpace()
x = torch.from_numpy(clouds.filled_box(0x5EED0001, n)).cuda()
qv = torch.from_numpy(clouds.filled_box(0x5EED0002, n)).cuda()
sp = torch.cat([qv, torch.full((n, 1), float(clouds.bvh_driver_radius(10)), device="cuda")], 1).contiguous()
bvh = abx.BoundingVolumeHierarchy(space, x)
for it in range(2):
    if knn:
        idx, off = bvh.query(space, abx.nearest(qv, 10))
    else:
        idx, off = bvh.query(space, abx.intersects(sp))
    torch.cuda.synchronize()
print("done", idx.numel())
