#!/usr/bin/env python
"""Which rays of the triangles_20M workload get different hit sets from the CUDA path and the oracle?"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import arborx_b200 as abx
import oracle
from tests import clouds

ref = int(os.environ.get("REFINEMENTS", "10"))
vert, tri = clouds.icosphere(ref)
soup = clouds.triangle_soup(vert, tri)
V = vert.shape[0]
rays = clouds.ball_rays(0x5EED0053, V)
cpu_q = 200_000
space = abx.ExecutionSpace()
bvh = abx.BoundingVolumeHierarchy(space, torch.from_numpy(soup).cuda(), abx.TRIANGLE)
# (a) the sample alone, (b) the sample as the head of the full batch
idx_a, off_a = bvh.query(space, abx.intersects(torch.from_numpy(rays[:cpu_q]).cuda(), abx.RAY_PRED))
idx_b, off_b = bvh.query(space, abx.intersects(torch.from_numpy(rays).cuda(), abx.RAY_PRED))
space.fence()
otree = oracle.Tree(soup, oracle.PRIM_TRI)
roff, ridx = otree.spatial_crs(rays[:cpu_q], oracle.PRED_RAY, True, 0)
off_a = off_a.cpu().numpy(); idx_a = idx_a.cpu().numpy()
off_b = off_b[:cpu_q + 1].cpu().numpy(); idx_b = idx_b[:int(off_b[-1])].cpu().numpy()
ca, cb, cr = np.diff(off_a), np.diff(off_b), np.diff(roff)
print("sample alone == oracle:", np.array_equal(ca, cr), " head of full batch == oracle:", np.array_equal(cb, cr),
      " alone == head:", np.array_equal(ca, cb))
bad = np.nonzero(cb != cr)[0]
print("rays with different counts:", len(bad), bad[:10])
out = []
for i in bad[:8]:
    g = sorted(idx_b[off_b[i]:off_b[i + 1]].tolist())
    r = sorted(ridx[roff[i]:roff[i + 1]].tolist())
    print(i, "gpu", g, "oracle", r, "ray", rays[i].tolist())
    out.append({"ray": [float(x) for x in rays[i]], "gpu": g, "oracle": r,
                "tris": {str(t): [float(x) for x in soup[t]] for t in set(g) | set(r)}})
# same sets for rays with equal counts?
diffsets = 0
for i in range(0, cpu_q, 97):
    if cb[i] == cr[i] and sorted(idx_b[off_b[i]:off_b[i + 1]].tolist()) != sorted(ridx[roff[i]:roff[i + 1]].tolist()):
        diffsets += 1
print("sampled rays with equal counts but different sets:", diffsets)
json.dump(out, open("gpurun_out/r02_debug_rays.json", "w"))
