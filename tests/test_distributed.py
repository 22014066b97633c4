"""DistributedTree: the reference-shaped exchange protocol on CPU (world sizes 1-3, gloo, oracle as the local
engine) and, with -m gpu, the C++ DistributedTree of libabx.so: over a one-rank NCCL communicator and as 2-4 ranks
(host threads) on one GPU over the in-process communicator; across GPUs via torchrun in scripts/dist_check.py."""
import os

import numpy as np
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
    """The C++ DistributedTree over a real (one-rank) NCCL communicator bootstrapped from the process group."""
    import arborx_b200 as abx
    from arborx_b200.distributed import DistributedTree
    from tests.distributed_cases import run_cases
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29555")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        space = abx.ExecutionSpace()
        make_tree = lambda v, kind=None: DistributedTree(dist.group.WORLD, space, v, kind)
        run_cases(0, 1, make_tree, torch.device("cuda", 0), space, check_host=True)
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 4])
def test_distributed_native_local_ranks(world):
    """The C++ DistributedTree with `world` ranks on ONE GPU: every rank is a host thread with its own
    execution space, the communicator is the in-process group (abx_comm_create_local), so routing, both
    exchanges, the sort by query id and the merge kernels run exactly as they do across GPUs."""
    import threading

    import arborx_b200 as abx
    from arborx_b200.distributed import Communicator, DistributedTree
    from tests.distributed_cases import run_cases
    comms = Communicator.local_group(world)
    errors = [None] * world

    def worker(r):
        try:
            torch.cuda.set_device(0)
            space = abx.ExecutionSpace(torch.cuda.Stream())
            with torch.cuda.stream(space.stream):
                make_tree = lambda v, kind=None: DistributedTree(comms[r], space, v, kind)
                run_cases(r, world, make_tree, torch.device("cuda", 0), space, check_host=True)
        except BaseException:
            import traceback
            errors[r] = traceback.format_exc()

    threads = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a collective: %s" % errors
    for r, e in enumerate(errors):
        assert e is None, "rank %d:\n%s" % (r, e)


@pytest.mark.gpu
@pytest.mark.parametrize("world,layout", [(2, "slabs"), (4, "blocks"), (3, "lopsided")])
def test_distributed_native_knn_large(world, layout):
    """kNN with a few 10^5 points and queries per rank, a rank with fewer than k points, and queries far outside
    every block (their k-th local distance reaches every rank).  Every rank's rows must be the rows of ONE tree
    over all points: same distances bit for bit, and every (index, rank) pair names a point at that distance.
    With ABX_LIBRARY = the tuning library and ABX_KNN_TWO_STAGE=1 (test_distributed_native_knn_two_stage_form) the
    same cases run the two-stage form of abx_dist.cu: near-boundary points first, their exchange under the interior
    points' traversal, a second exchange for the far-away queries."""
    import threading

    import arborx_b200 as abx
    from arborx_b200.distributed import Communicator, DistributedTree
    from tests import clouds
    n_per, q_per, k = 300_000, 280_000, 7
    rng = np.random.default_rng(1234)
    pts, qs = [], []
    for r in range(world):
        p = clouds.filled_box(0xABC0 + r, n_per).astype(np.float32)
        a = float(np.cbrt(n_per))
        p = p / a * 0.5 + 0.5  # unit cube
        if layout == "slabs":
            off = np.array([r, 0, 0], np.float32)
        elif layout == "blocks":
            off = np.array([r % 2, r // 2, 0], np.float32)
        else:
            off = np.array([r, 0, 0], np.float32)
            if r == 1:
                p = p[:5]  # fewer than k points on this rank, and no split there
        p = p + off
        qq = rng.random((q_per, 3), dtype=np.float32) + off
        qq[:2000] = rng.random((2000, 3), dtype=np.float32) * 40 - 20  # far outside: interior by the guess, remote by k-th distance
        m = min(100, len(p))
        qq[2000:2000 + m] = p[:m]
        pts.append(np.ascontiguousarray(p))
        qs.append(np.ascontiguousarray(qq))
    all_pts = np.concatenate(pts)
    starts = np.cumsum([0] + [len(p) for p in pts])
    space0 = abx.ExecutionSpace()
    whole = abx.BoundingVolumeHierarchy(space0, torch.from_numpy(all_pts).cuda())
    expected = []
    for r in range(world):
        idx, off, d = whole.query(space0, abx.nearest(torch.from_numpy(qs[r]).cuda(), k), return_distances=True)
        expected.append(d.view(-1, k).cpu().numpy())
    comms = Communicator.local_group(world)
    errors = [None] * world

    def worker(r):
        try:
            torch.cuda.set_device(0)
            space = abx.ExecutionSpace(torch.cuda.Stream())
            with torch.cuda.stream(space.stream):
                tree = DistributedTree(comms[r], space, torch.from_numpy(pts[r]).cuda())
                for rep in range(2):
                    vals, off, d = tree.query(space, abx.nearest(torch.from_numpy(qs[r]).cuda(), k),
                                              return_distances=True)
                    space.fence()
                    off = off.cpu().numpy()
                    assert np.array_equal(off, np.arange(q_per + 1) * k)
                    d = d.view(-1, k).cpu().numpy()
                    assert np.array_equal(d, expected[r]), "rank %d rep %d: distances differ" % (r, rep)
                    v = vals.cpu().numpy().reshape(-1, k, 2)
                    g = starts[v[..., 1]] + v[..., 0]
                    assert (v[..., 1] >= 0).all() and (v[..., 1] < world).all()
                    diff = all_pts[g] - qs[r][:, None, :]
                    dd = np.sqrt((diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]).astype(np.float32)
                                 + diff[..., 2] * diff[..., 2]).astype(np.float32)
                    assert np.array_equal(dd, d), "rank %d: a pair does not name a point at its distance" % r
                    assert all(len(set(row)) == k for row in g[:5000].tolist())
                # host form (compact results)
                hv, hoff, hd, rpos, rrank = tree.query(space, abx.nearest(torch.from_numpy(qs[r]), k),
                                                       return_distances=True)
                assert np.array_equal(hd.view(-1, k).numpy(), expected[r])
                owner = np.full(hv.numel(), r, np.int64)
                owner[rpos.numpy()] = rrank.numpy()
                assert np.array_equal(starts[owner] + hv.numpy(), g.reshape(-1))
        except BaseException:
            import traceback
            errors[r] = traceback.format_exc()

    threads = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a collective: %s" % errors
    for r, e in enumerate(errors):
        assert e is None, "rank %d:\n%s" % (r, e)


@pytest.mark.gpu
def test_distributed_native_knn_two_stage_form():
    """The measured-and-shelved two-stage kNN of abx_dist.cu stays correct: same cases, tuning library."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tuning = os.path.join(root, "arborx_b200", "lib", "libabx_tuning.so")
    if not os.path.exists(tuning):
        pytest.skip("libabx_tuning.so not built (make -C arborx_b200/csrc tuning)")
    env = dict(os.environ, ABX_LIBRARY=tuning, ABX_KNN_TWO_STAGE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.abspath(__file__), "-k",
                        "knn_large"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


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
    """abx_dist_dbscan_points3f over a one-rank NCCL communicator."""
    import arborx_b200 as abx
    from arborx_b200.distributed_dbscan import dbscan as dist_dbscan
    from tests.distributed_dbscan_cases import run_cases
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29556")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        space = abx.ExecutionSpace()
        run = lambda pts, eps, minpts, params: dist_dbscan(dist.group.WORLD, space, pts, eps, minpts, params)
        assert run_cases(0, 1, run, lambda a: [a], torch.device("cuda", 0), n=20000) == 8
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 4])
def test_distributed_dbscan_native_local_ranks(world):
    """The C++ distributed DBSCAN with `world` ranks (host threads) on one GPU over the in-process communicator:
    halo exchange, ghost labels back, merge pairs, all-gather of the pairs and relabelling as across GPUs."""
    import threading

    import arborx_b200 as abx
    from arborx_b200.distributed import Communicator
    from arborx_b200.distributed_dbscan import dbscan as dist_dbscan
    from tests.distributed_dbscan_cases import run_cases
    comms = Communicator.local_group(world)
    errors = [None] * world
    barrier = threading.Barrier(world)
    shared = {}

    def worker(r):
        try:
            torch.cuda.set_device(0)
            space = abx.ExecutionSpace(torch.cuda.Stream())

            def gather(a):
                shared[r] = a
                barrier.wait(timeout=120)
                out = [shared[k] for k in range(world)]
                barrier.wait(timeout=120)
                return out

            with torch.cuda.stream(space.stream):
                run = lambda pts, eps, minpts, params: dist_dbscan(comms[r], space, pts, eps, minpts, params)
                assert run_cases(r, world, run, gather, torch.device("cuda", 0), n=6000) == 8
        except BaseException:
            import traceback
            errors[r] = traceback.format_exc()
            barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a collective: %s" % errors
    for r, e in enumerate(errors):
        assert e is None, "rank %d:\n%s" % (r, e)
