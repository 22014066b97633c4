"""ABX_DIST_DEBUG=1 torchrun --nproc-per-node N scripts/dist_timeline.py : section timeline of the distributed step"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx
from arborx_b200.distributed import DistributedTree
import bench
lr = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rank, world = dist.get_rank(), dist.get_world_size()
n = 10_000_000
values, queries, spheres, r = bench.make_inputs(n, n, rank, world)
space = abx.ExecutionSpace()
dv, dq, ds = (torch.from_numpy(a).cuda() for a in (values, queries, spheres))
for it in range(4):
    tree = DistributedTree(dist.group.WORLD, space, dv)
    v, o = tree.query(space, abx.intersects(ds))
    v2, o2 = tree.query(space, abx.nearest(dq, 10))
dist.barrier()
dist.destroy_process_group()
