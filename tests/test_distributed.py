"""DistributedTree: the N>1 exchange protocol on CPU (world_size-2/3 gloo, oracle as the local engine) and
on the GPU (-m gpu: single-rank nccl group in-process; multi-rank via torchrun in scripts/dist_check.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from tests.distributed_cases import OracleEngine, run_all
        run_all(OracleEngine, torch.device("cpu"))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_distributed_protocol_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", "rank %d:\n%s" % (r, msg)


@pytest.mark.gpu
def test_distributed_single_rank_cuda():
    import arborx_b200 as abx
    from arborx_b200.distributed import CudaEngine
    from tests.distributed_cases import run_all
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29555")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        space = abx.ExecutionSpace()
        run_all(lambda: CudaEngine(space), torch.device("cuda", 0), space)
    finally:
        if created:
            dist.destroy_process_group()


def _dbscan_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from tests.distributed_dbscan_cases import OracleDBSCANEngine, run_all
        assert run_all(OracleDBSCANEngine, torch.device("cpu")) == 8
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("world", [1, 2, 3])
def test_distributed_dbscan_gloo(world):
    """Distributed DBSCAN halo exchange / label merge on CPU (oracle as the local engine)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_dbscan_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", "rank %d:\n%s" % (r, msg)


@pytest.mark.gpu
def test_distributed_dbscan_single_rank_cuda():
    import arborx_b200 as abx
    from arborx_b200.distributed_dbscan import CudaDBSCANEngine
    from tests.distributed_dbscan_cases import run_all
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29556")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        space = abx.ExecutionSpace()
        assert run_all(lambda s: CudaDBSCANEngine(s), torch.device("cuda", 0), space, n=20000) == 8
    finally:
        if created:
            dist.destroy_process_group()
