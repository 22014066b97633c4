"""Three builds at n points for ncu (build kernels only):
    ncu --set full --clock-control none --import-source on -k regex:'hierarchy|segmentFix' -s 4 -c 3 \
        -o gpurun_out/prof_build python scripts/profile_build.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
space = abx.ExecutionSpace()
x = torch.from_numpy(clouds.filled_box(0x5EED0001, n)).cuda()
for it in range(3):
    bvh = abx.BoundingVolumeHierarchy(space, x)
    torch.cuda.synchronize()
print("done", bvh.size())
