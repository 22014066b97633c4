"""Multi-GPU DistributedTree + distributed DBSCAN check: torchrun --nproc-per-node N scripts/dist_check.py
The C++ DistributedTree (libabx.so over NCCL) runs the rank-count-agnostic cases of tests/distributed_cases.py
(device results and the host / compact entry points), then a larger cross-check against the torch protocol
model with the CUDA engine, then the distributed DBSCAN cases."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx
from arborx_b200.distributed import CudaEngine, DistributedTree
from tests.distributed_cases import run_cases, rows, _Pred
from tests import clouds
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
space = abx.ExecutionSpace()
comm = dist.group.WORLD
run_cases(rank, world, lambda v, kind=None: DistributedTree(comm, space, v, kind), dev, space, check_host=True)
# larger clouds: native tree vs the protocol model on the same data
n, q = 200_000, 50_000
pts = torch.from_numpy(clouds.uniform01(100 + rank, n) + np.float32(0.8 * rank) * np.array([1, 0, 0], np.float32)).to(dev)
qs = torch.from_numpy(clouds.uniform01(900 + rank, q) * np.array([0.8 * (world - 1) + 1, 1, 1], np.float32)).to(dev)
native = DistributedTree(comm, space, pts)
model = DistributedTree(comm, space, pts, engine=CudaEngine(space))
sp = torch.cat([qs, torch.full((q, 1), 0.02, device=dev)], 1).contiguous()
for pred, kw in ((_Pred("spatial", 0, sp), {}), (_Pred("nearest", 2, qs, 8), {"return_distances": True})):
    a = native.query(space, pred, **kw)
    b = model.query(space, pred, **kw)
    assert torch.equal(a[1].cpu(), b[1].cpu()), "offsets differ"
    if pred.tag == "nearest":
        assert torch.equal(a[2].cpu(), b[2].cpu()), "distances differ"
    else:
        assert rows(a[0], a[1]) == rows(b[0], b[1]), "rows differ"
# distributed DBSCAN: the C++ path (abx_dist_dbscan_points3f) and the torch protocol model with the CUDA engine
from arborx_b200.distributed_dbscan import CudaDBSCANEngine, dbscan as dist_dbscan
from tests.distributed_dbscan_cases import run_all as run_dbscan_model, run_cases as run_dbscan_cases


def gather(a):
    out = [None] * world
    dist.all_gather_object(out, a)
    return out


assert run_dbscan_cases(rank, world, lambda p, eps, m, prm: dist_dbscan(comm, space, p, eps, m, prm), gather, dev, n=40000) == 8
assert run_dbscan_model(lambda s: CudaDBSCANEngine(s), dev, space, n=40000) == 8
dist.barrier()
if rank == 0:
    print("DIST CHECK OK world=%d" % world)
dist.destroy_process_group()
