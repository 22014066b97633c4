"""Two DBSCAN runs on a GanTao cloud for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'denseMainKernel|fdbscanMainKernel' -s 1 -c 1 \
        -o gpurun_out/prof_dbscan python scripts/profile_dbscan.py [n] [impl]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
space = abx.ExecutionSpace()
d = torch.from_numpy(clouds.gan_tao(3, n)).cuda()
for it in range(2):
    labels = abx.dbscan(space, d, 200.0, 5, abx.DBSCANParameters(impl, 0))
    torch.cuda.synchronize()
print("done", int((labels < 0).sum()))
