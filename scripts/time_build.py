"""Build-only timing at n points (CUDA events around abx_bvh_build, per-kernel table):
    ABX_HIER_WARPS=4 python scripts/time_build.py [n] [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kind = sys.argv[3] if len(sys.argv) > 3 else "points"
space = abx.ExecutionSpace()
pts = clouds.filled_box(0x5EED0001, n)
if kind == "boxes":
    import numpy as np
    pts = np.concatenate([pts, pts + np.float32(0.5)], 1)
x = torch.from_numpy(pts).cuda()
for it in range(3):
    bvh = abx.BoundingVolumeHierarchy(space, x)
torch.cuda.synchronize()
abx.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for it in range(reps):
    e0.record()
    bvh = abx.BoundingVolumeHierarchy(space, x)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print("build n=%d %s: median %.4f ms min %.4f ms  (%.2f Gprims/s)" % (n, kind, ts[len(ts) // 2], ts[0], n / ts[len(ts) // 2] / 1e6))
for name, cnt, ms, _mx in abx.profile_report():
    print("   %-40s %4d  %.4f ms/launch  %.4f ms/build" % (name, cnt, ms / cnt, ms / reps))
