"""Host-side timeline of the radius CRS query (where does non-kernel time go?)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx
from tests import clouds
n = 10_000_000
space = abx.ExecutionSpace()
x = torch.from_numpy(clouds.filled_box(0x5EED0001, n)).cuda()
qv = torch.from_numpy(clouds.filled_box(0x5EED0002, n)).cuda()
sp = torch.cat([qv, torch.full((n, 1), float(clouds.bvh_driver_radius(10)), device="cuda")], 1).contiguous()
bvh = abx.BoundingVolumeHierarchy(space, x)
p = abx.intersects(sp)
pn = abx.nearest(qv, 10)
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, off = bvh.query(space, p)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    kidx, koff = bvh.query(space, pn)
    t3 = time.perf_counter(); torch.cuda.synchronize(); t4 = time.perf_counter()
    b2 = abx.BoundingVolumeHierarchy(space, x)
    t5 = time.perf_counter(); torch.cuda.synchronize(); t6 = time.perf_counter()
    print("it%d radius: call %.2f ms (+sync %.2f) | knn: call %.2f (+sync %.2f) | build: call %.2f (+sync %.2f) | reserved %.0f MB"
          % (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t6-t5)*1e3, torch.cuda.memory_reserved()/1e6), flush=True)
    del idx, off, kidx, koff, b2
