"""Two identical iterations of the bench step (build + radius CRS + kNN) at 10M for
ncu: skip the kernels of the first iteration, capture those of the second.
    ncu --set full --clock-control none --import-source on \
        -k regex:'onesweepPassKernel|hierarchyKernel|spatialKernel|nearestKernel' -s 20 -c 20 \
        -o gpurun_out/prof python scripts/profile_kernels.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
space = abx.ExecutionSpace()
x = torch.from_numpy(clouds.filled_box(0x5EED0001, n)).cuda()
qv = torch.from_numpy(clouds.filled_box(0x5EED0002, n)).cuda()
sp = torch.cat([qv, torch.full((n, 1), float(clouds.bvh_driver_radius(10)), device="cuda")], 1).contiguous()
for it in range(2):
    bvh = abx.BoundingVolumeHierarchy(space, x)
    idx, off = bvh.query(space, abx.intersects(sp))
    kidx, koff = bvh.query(space, abx.nearest(qv, 10))
    torch.cuda.synchronize()
print("done", idx.numel(), kidx.numel())
