"""One MinimumSpanningTree run on a GanTao cloud for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'componentNearestKernel|reduceLabelsKernel' \
        -s 2 -c 6 -o gpurun_out/prof_mst python scripts/profile_mst.py [n] [k]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
space = abx.ExecutionSpace()
d = torch.from_numpy(clouds.gan_tao(3, n)).cuda()
mst = abx.MinimumSpanningTree(space, d, k)
torch.cuda.synchronize()
print("done", mst.iterations, float(mst.weights.double().sum()))
