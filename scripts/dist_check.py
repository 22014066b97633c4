"""Multi-GPU DistributedTree + distributed DBSCAN check: torchrun --nproc-per-node N scripts/dist_check.py"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx
from arborx_b200.distributed import CudaEngine
from tests.distributed_cases import run_all
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
space = abx.ExecutionSpace()
run_all(lambda: CudaEngine(space), torch.device("cuda", lr), space)
from arborx_b200.distributed_dbscan import CudaDBSCANEngine
from tests.distributed_dbscan_cases import run_all as run_dbscan
assert run_dbscan(lambda s: CudaDBSCANEngine(s), torch.device("cuda", lr), space, n=40000) == 8
dist.barrier()
if dist.get_rank() == 0:
    print("DIST CHECK OK world=%d" % dist.get_world_size())
dist.destroy_process_group()
