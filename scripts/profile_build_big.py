"""Per-kernel times of BVH builds at several sizes (where does the 100M build go?):
    python scripts/profile_build_big.py 10000000 50000000 100000000"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

space = abx.ExecutionSpace()
for n in [int(a) for a in sys.argv[1:]] or [10_000_000, 100_000_000]:
    x = clouds.filled_box_torch(0x5EED0101, n, "cuda")
    for it in range(3):
        if it == 2:
            abx.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bvh = abx.BoundingVolumeHierarchy(space, x)
        e1.record()
        torch.cuda.synchronize()
        del bvh
    prof = abx.profile_report()
    abx.profile_enable(False)
    print("n = %d: build %.3f ms (%.2f Gprims/s)" % (n, e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6))
    for name, cnt, ms, mx in prof[:10]:
        print("    %-40s %3d  %8.3f ms" % (name, cnt, ms))
    del x
    abx.trim()
