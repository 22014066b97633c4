#!/usr/bin/env python
"""Phase timeline of the distributed queries (torchrun, tuning library): ABX_LIBRARY=.../libabx_tuning.so
ABX_DIST_TRACE=1.  Every phase is drained before the next one starts, so the lines attribute the time instead of
reproducing the overlapped call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import arborx_b200 as abx
import bench
from arborx_b200.distributed import Communicator, DistributedTree


def main():
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = q = int(os.environ.get("TRACE_N", "10000000"))
    values, queries, spheres, r = bench.make_inputs(n, q, rank, world)
    space = abx.ExecutionSpace()
    d_values = torch.from_numpy(values).cuda()
    p_spatial = abx.intersects(torch.from_numpy(spheres).cuda())
    p_nearest = abx.nearest(torch.from_numpy(queries).cuda(), bench.K_NEIGHBORS)
    comm = Communicator.from_process_group(dist.group.WORLD)
    for it in range(int(os.environ.get("TRACE_STEPS", "4"))):
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            sys.stderr.write("--- step %d\n" % it)
        tree = DistributedTree(comm, space, d_values)
        tree.query(space, p_spatial)
        tree.query(space, p_nearest)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
