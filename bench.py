#!/usr/bin/env python
"""bench.py -- the bvh_driver workload of BASELINE.json (configs[1]) on B200, plus the other
BASELINE configs as extra blocks of the same JSON line.

One "step" (the headline, timed region) = one pass of the hot path over one batch of synthetic input:
    BVH build over n points  +  intersects(sphere) CRS query with q spheres
    +  nearest(k=10) CRS query with q points
with n = q = 10M filled-box points, r = cbrt(10*6/pi), predicates Morton-sorted,
buffer_size = 0 (benchmarks/bvh_driver/benchmark_registration.hpp:129-205).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

prints ONE JSON line (see the task contract).  `value` is the aggregate item rate
(n + 2q items per step) with inputs resident in HBM; `components` holds the three
rates BASELINE.json names (build Mprims/s, radius Mqueries/s, kNN Mqueries/s) with their roofline
fractions; `e2e` is the same step through the host-buffer C-ABI entry points (pinned host inputs,
results copied back to the host).  Outside the timed region, `workloads` adds:
    dbscan_10M        BASELINE configs[2]: ArborX::dbscan on the GanTao cloud (FDBSCAN / DenseBox, minpts 2 / 5)
    triangles_20M     BASELINE configs[4]: BVH over the 20 971 520-triangle icosphere, nearest(point, 1), rays   (N = 1)
    distributed_100M  BASELINE configs[3]: DistributedTree over 100M points in total, sharded over the N GPUs (strong)
    distributed_dbscan  configs[2] at N > 1 (weak scaling, halo exchange + label merge)
Roofline figures come in two kinds: `model_frac` = SURVEY 8(d) algorithmic bytes / time / peak (the reference
algorithm's node-visit traffic: an upper estimate of HBM traffic, cache hits included -- it can exceed 1) and
`dram_frac` = DRAM bytes of the kernel measured by ncu (profiles/traffic.json, committed capture) / time / peak.

--impl reference times the CPU restatement of the reference (oracle/, OpenMP, all host threads; the real
reference cannot be built here: Kokkos absent) on a bounded sample of the same workload per step.
Multi-GPU (--gpus N under torchrun): the path shards in DistributedTree -- every rank owns n points / q queries
of a touching block lattice (distributed_tree_driver layout), the step is DistributedTree build + distributed
radius + distributed kNN through the C++ DistributedTree (libabx.so over NCCL), weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout must carry exactly ONE JSON line.  Libraries write banners to fd 1 (NCCL prints its
# version there), so fd 1 is pointed at stderr for the whole run and the result line is written
# to the saved descriptor by emit().
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


METRIC = "BVH build Mprims/s; radius & kNN(k=10) Mqueries/s at 10M pts; % HBM roofline"
UNIT = "Mitems/s (n prims + q radius queries + q kNN queries per step second)"
K_NEIGHBORS = 10


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def traffic_table():
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(tp)) if os.path.exists(tp) else {}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def block_offset(n, rank, world):
    """distributed_tree_driver layout (benchmarks/distributed_tree_driver/distributed_tree_driver.cpp:44-149):
    blocks of side 2a, a = cbrt(n), on a ceil(cbrt(R))^3 lattice, shift = 1 (touching)."""
    nb = int(np.ceil(np.cbrt(world) - 1e-9))
    ijk = np.array([rank % nb, (rank // nb) % nb, rank // (nb * nb)], np.float32)
    a = np.float32(np.cbrt(float(n)))
    return (np.float32(2) * a * ijk).astype(np.float32)


def make_inputs(n, q, rank=None, world=1):
    """rank is None: the single-tree bvh_driver clouds.  Otherwise this rank's block of the
    distributed_tree_driver layout."""
    from tests import clouds
    if rank is None:
        values = clouds.filled_box(0x5EED0001, n)
        queries = clouds.filled_box(0x5EED0002, q)
    else:
        off = block_offset(n, rank, world)
        values = (clouds.filled_box(0x5EED0001 + 16 * rank, n) + off).astype(np.float32)
        queries = (clouds.filled_box(0x5EED0002 + 16 * rank, q) + off).astype(np.float32)
    r = clouds.bvh_driver_radius(K_NEIGHBORS)
    spheres = np.concatenate([queries, np.full((q, 1), r, np.float32)], 1).astype(np.float32)
    return values, queries, spheres, float(r)


def workload_config(n, q, r, world=1):
    cfg = {"workload": "bvh_driver filled_box: build + intersects(sphere) CRS + nearest(k=10) CRS "
                       "(BASELINE.json configs[1])",
           "n_values": n, "n_queries": q, "k": K_NEIGHBORS, "radius": r, "sort_predicates": True, "buffer_size": 0,
           "cloud": "uniform in [-cbrt(n), cbrt(n)]^3, counter-based RNG (tests/clouds.py)",
           "cache": "inputs and tree (%.0f MB nodes) larger than the 126 MB L2; no explicit flush" % (64e-6 * n)}
    if world > 1:
        cfg["parallelism"] = ("DistributedTree over %d GPUs: %d points and %d queries per rank on a touching block "
                              "lattice; routing kernel + grouped NCCL send/recv (libabx.so)" % (world, n, q))
        cfg["workload"] += " through DistributedTree (BASELINE.json configs[3] layout, weak scaling)"
    return cfg


# ------------------------------------------------------------------ CPU arm ----
def oracle_all_threads():
    """torchrun exports OMP_NUM_THREADS=1: the CPU arm always asks for every host thread it may use."""
    import oracle
    oracle.lib().orc_set_num_threads(host_threads())
    return oracle


def cpu_sample_inputs(n_s, world):
    """The bounded sample of the step the CPU arm times: the same workload at n_s points / n_s queries (same
    density, radius and k: the 1M-point case of BASELINE configs[0] by default); at N > 1 the union of the N
    rank blocks of the distributed layout, n_s / N points each, as ONE shared-memory tree on this host."""
    if world == 1:
        return make_inputs(n_s, n_s)
    per = n_s // world
    parts = [make_inputs(per, per, r, world) for r in range(world)]
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
            np.concatenate([p[2] for p in parts]), parts[0][3])


def cpu_step(oracle, values, queries, spheres):
    t0 = time.perf_counter()
    tree = oracle.Tree(values)
    t1 = time.perf_counter()
    off, idx = tree.spatial_crs(spheres, oracle.PRED_SPHERE, True, 0)
    t2 = time.perf_counter()
    koff, kidx, kd = tree.nearest_crs(queries, K_NEIGHBORS, True)
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2, int(off[-1])


def cpu_counters(oracle, values, queries, spheres, q_sample):
    """Traversal counters of the reference algorithm (reference-layout tree, rope / stack traversal) for the
    byte model of SURVEY 8(d): internal-node box tests I and leaf tests L per query."""
    tree = oracle.Tree(values)
    _, c_sp = tree.spatial_count(spheres[:q_sample], counters=True)
    _, _, _, c_nn = tree.nearest_crs(queries[:q_sample], K_NEIGHBORS, True, counters=True)
    return {"spatial_I": float(c_sp[0]) / q_sample, "spatial_L": float(c_sp[1]) / q_sample,
            "nearest_I": float(c_nn[0]) / q_sample, "nearest_L": float(c_nn[1]) / q_sample, "q_sample": q_sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle = oracle_all_threads()
    world = max(1, args.gpus)
    n, q = args.n, args.q
    n_s = min(n, args.cpu_n)
    values, queries, spheres, r = cpu_sample_inputs(n_s, world)
    n_s = len(values)
    times = []
    for it in range(args.warmup + args.steps):
        tb, tr, tk, _ = cpu_step(oracle, values, queries, spheres)
        if it >= args.warmup:
            times.append((tb + tr + tk, tb, tr, tk))
    t = np.mean(np.array(times), 0)
    value = 3 * n_s / t[0] / 1e6
    sample = ("every step = the whole workload at 1/%d scale: build of %d points + %d radius + %d kNN queries "
              "(same density, radius, k; %s), measured, not extrapolated; oracle/arborx_oracle.cpp, OpenMP"
              % (max(1, n // max(n_s // world, 1)), n_s, n_s, n_s,
                 "one tree over the union of the %d rank blocks" % world if world > 1 else "BASELINE configs[0] size"))
    cfg = workload_config(n, q, r, world)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t[0] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "components": {"build_Mprims_s": n_s / t[1] / 1e6, "radius_Mqueries_s": n_s / t[2] / 1e6,
                       "knn_Mqueries_s": n_s / t[3] / 1e6},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample,
                         "extrapolated": False},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------ byte models ----
def traversal_bytes(I, L, leaf_bytes=20):
    return 32.0 * I + leaf_bytes * L


def kernel_models(ctr, n, q, nnz):
    """SURVEY 8(d) algorithmic bytes per launch of the kernels of the headline step."""
    sp = traversal_bytes(ctr["spatial_I"], ctr["spatial_L"])
    nn = traversal_bytes(ctr["nearest_I"], ctr["nearest_L"])
    return {
        "spatialKernel<count>": q * (sp + 16 + 4),
        "spatialKernel<fill>": q * (sp + 16 + 4) + 4.0 * nnz,
        "spatialKernel<stage>": q * (sp + 16 + 4) + 4.0 * nnz,
        "spatialKernel<compact>": 8.0 * q + 12.0 * nnz,
        "hierarchyLocalKernel": 64.0 * n + 36.0 * n,
        "hierarchyGlobalKernel": 0.05 * n * (36 + 64 + 64),
        "nearestKernel": q * (nn + 12 + 4 * K_NEIGHBORS),
        "onesweepPassKernel<u64>": 24.0 * n,
        "onesweepPassKernel<u32>": 16.0 * q,
        "segmentFixKernel": 24.0 * n,
        "radixHistogramKernel<u64>": 8.0 * n,
        "morton64Kernel": 20.0 * n,
        "sceneBoundsKernel": 12.0 * n,
    }


def kernel_table(prof, models, traffic, peak, launches_per_step_hint=1):
    """Per kernel: launches, total, average and LONGEST launch; model / DRAM rates for the longest launch (a
    step may hold a full-size launch plus small ones for forwarded queries: averaging them would halve the
    launch time and double the rate)."""
    rows = []
    for name, cnt, ms, mx in prof:
        key = None
        for kname in models:
            if kname in name or ("<" not in kname and kname in name.split("<")[0]):
                key = kname
        row = {"kernel": name, "launches": cnt, "total_ms": round(ms, 4), "avg_ms": round(ms / cnt, 5),
               "max_ms": round(mx, 5)}
        if key:
            row["model_bytes"] = models[key]
            row["model_gbs"] = models[key] / (mx * 1e-3) / 1e9
            row["model_frac"] = row["model_gbs"] / peak
        base = name.split("<")[0].strip("( ")
        t = traffic.get(name, traffic.get(base))
        if t:
            row["dram_bytes_ncu"] = t
            row["dram_gbs"] = t / (mx * 1e-3) / 1e9
            row["dram_frac"] = row["dram_gbs"] / peak
        rows.append(row)
    return rows


def phase_fracs(name, model_bytes, dram_bytes, ms, peak):
    out = {name + "_model_bytes": model_bytes, name + "_model_frac": model_bytes / (ms * 1e-3) / 1e9 / peak}
    if out[name + "_model_frac"] > 1.0:
        # the model counts the reference algorithm's node visits (two traversal passes for the radius phase, cache
        # hits included): above the HBM peak it is a statement about the model, not a roofline
        out[name + "_model_exceeds_peak"] = True
    if dram_bytes:
        out[name + "_dram_bytes_ncu"] = dram_bytes
        out[name + "_dram_frac"] = dram_bytes / (ms * 1e-3) / 1e9 / peak
    return out


# ------------------------------------------------------------------ GPU arm ----
def pin_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs closest to its GPU before any pinned host memory is allocated, so that the
    end-to-end copies of the N ranks do not all cross one socket's memory controller."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            # the CUDA ordinal is not the NVML index when CUDA_VISIBLE_DEVICES is set: go through the PCI address
            p = torch.cuda.get_device_properties(local_rank)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(
                ("%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)).encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


def run_ours(args):
    import torch
    import torch.distributed as dist

    import arborx_b200 as abx
    from arborx_b200.distributed import Communicator, DistributedTree

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    # host threads next to the GPU for the device and end-to-end legs (pinned buffers are allocated from here on);
    # the full mask comes back before any CPU-arm work so that the OpenMP restatement sees every core
    full_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_pinned = False if os.environ.get("ABX_BENCH_NO_PIN") else pin_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, q = args.n, args.q
    values, queries, spheres, r = make_inputs(n, q, rank if world > 1 else None, world)
    space = abx.ExecutionSpace()
    d_values = torch.from_numpy(values).cuda()
    d_queries = torch.from_numpy(queries).cuda()
    d_spheres = torch.from_numpy(spheres).cuda()
    p_spatial = abx.intersects(d_spheres)
    p_nearest = abx.nearest(d_queries, K_NEIGHBORS)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    comm = Communicator.from_process_group(dist.group.WORLD) if world > 1 else None

    def make_tree(vals):
        # N > 1: the path shards in DistributedTree (primitives per GPU, replicated top tree, grouped NCCL
        # send/recv of forwarded queries and results, all inside libabx.so); N = 1: the single tree
        return DistributedTree(comm, space, vals) if world > 1 else abx.BoundingVolumeHierarchy(space, vals)

    def step(timers=None):
        e = [ev() for _ in range(4)] if timers is not None else None
        if e:
            e[0].record()
        bvh = make_tree(d_values)
        if e:
            e[1].record()
        idx, off = bvh.query(space, p_spatial)
        if e:
            e[2].record()
        kidx, koff = bvh.query(space, p_nearest)
        if e:
            e[3].record()
            timers.append(e)
        return idx.shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        nnz = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    abx.profile_enable(True)
    launches0 = abx.launch_count()
    timers = []
    t_start, t_stop = ev(), ev()
    barrier()
    t_start.record()
    for _ in range(args.steps):
        nnz = step(timers)
    t_stop.record()
    barrier()
    launches = abx.launch_count() - launches0
    prof = abx.profile_report()
    abx.profile_enable(False)
    clocks = sampler.stop()
    elapsed_ms = t_start.elapsed_time(t_stop)
    per_step = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in timers])
    if os.environ.get("ABX_BENCH_DEBUG"):
        log("per-step (build, radius, knn) ms:\n", per_step)
    parts = per_step.mean(0)  # build, radius, knn
    if world > 1:
        t = torch.tensor([elapsed_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * (n + 2 * q) / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI entry points -----------------------
    h_values = torch.from_numpy(values).pin_memory()
    h_spheres = torch.from_numpy(spheres).pin_memory()
    h_queries = torch.from_numpy(queries).pin_memory()
    import concurrent.futures
    E2E_THREADS, E2E_CHUNKS = args.e2e_threads, args.e2e_chunks
    e2e_pool = concurrent.futures.ThreadPoolExecutor(E2E_THREADS)
    e2e_local = threading.local()
    bounds = [q * c // E2E_CHUNKS for c in range(E2E_CHUNKS + 1)]
    hp_spatial_parts = [abx.intersects(h_spheres[bounds[c]:bounds[c + 1]]) for c in range(E2E_CHUNKS)]
    hp_nearest_parts = [abx.nearest(h_queries[bounds[c]:bounds[c + 1]], K_NEIGHBORS) for c in range(E2E_CHUNKS)]
    dist_pools = {"spatial": abx.HostBufferPool(), "nearest": abx.HostBufferPool()}
    part_pools = [abx.HostBufferPool() for _ in range(2 * E2E_CHUNKS)]

    def e2e_task(bvh, preds, slot):
        if not hasattr(e2e_local, "space"):
            torch.cuda.set_device(local_rank)
            e2e_local.space = abx.ExecutionSpace(torch.cuda.Stream())
        # results alias the part's own pool of pinned buffers (one task per part at a time): valid until that
        # part is queried again
        idx, off = bvh.query(e2e_local.space, preds, out=part_pools[slot])
        return int(off[-1]), idx.numel()

    def e2e_step():
        if world > 1:
            # abx_dist_create_host + abx_dist_query_spatial_crs_host + abx_dist_query_nearest_crs_host: host
            # primitives and predicates in, compact results (indices + the short list of entries owned by other
            # ranks) out to pinned host buffers; the calls are collective, so they are issued in program order
            tree = DistributedTree(comm, space, h_values)
            idx, off, rpos, rrank = tree.query(space, abx.intersects(h_spheres), out=dist_pools["spatial"])
            kidx, koff, kpos, krank = tree.query(space, abx.nearest(h_queries, K_NEIGHBORS), out=dist_pools["nearest"])
            return (int(off[-1]) + int(koff[-1]), idx.numel(), kidx.numel(),
                    rpos.numel() + rrank.numel() + kpos.numel() + krank.numel())
        bvh = abx.BoundingVolumeHierarchy(space, h_values)
        space.fence()
        # The query batches are independent, and so are their parts: each batch is issued as E2E_CHUNKS host
        # calls from E2E_THREADS host threads, every thread on its own execution space instance (stream), so
        # the result copy of one call overlaps the traversal of the next (PCIe is the long pole of this path:
        # 0.4 GB in, 0.88 GB out per step).  Spatial parts first: they have the most result bytes per traversal
        # millisecond, so their copies run under the kNN traversals (scripts/e2e_timeline.py,
        # profiles/r02_e2e_timeline.log: 26.6 ms against 30.5 ms with the kNN parts first).
        parts_all = hp_spatial_parts + hp_nearest_parts
        futures = [e2e_pool.submit(e2e_task, bvh, p, i) for i, p in enumerate(parts_all)]
        total, n_idx, n_kidx = 0, 0, 0
        for f, p in zip(futures, parts_all):
            last, cnt = f.result()
            total += last
            if p.tag == "nearest":
                n_kidx += cnt
            else:
                n_idx += cnt
        # results are host tensors: their last offsets were read on the host inside the tasks
        return total, n_idx, n_kidx, 0

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, n_idx, n_kidx, n_extra = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * (n + 2 * q) / e2e_s / 1e6
    h2d = 12 * n + 16 * q + 12 * q
    d2h = 4 * (q + 1) + 4 * n_idx + 4 * (q + 1) + 4 * n_kidx + 4 * n_extra
    del h_values, h_spheres, h_queries, hp_spatial_parts, hp_nearest_parts, dist_pools
    e2e_pool.shutdown()
    del d_values, d_queries, d_spheres, p_spatial, p_nearest
    abx.trim()

    if full_affinity is not None:
        os.sched_setaffinity(0, full_affinity)
    peak, peak_src, _ = peaks()
    traffic = traffic_table()

    # ---- the other BASELINE configs (outside the headline timed region) ---------------
    workloads = {}
    if not args.skip_workloads:
        for name, fn in (("bvh_1M", wl_bvh_1m), ("dbscan_10M", wl_dbscan), ("mst_10M", wl_mst),
                         ("triangles_20M", wl_triangles),
                         ("distributed_100M", wl_distributed), ("distributed_dbscan", wl_distributed_dbscan)):
            try:
                t0 = time.perf_counter()
                res = fn(args, abx, torch, dist, space, comm, rank, world, peak, traffic)
                if res is not None:
                    res["bench_wall_s"] = round(time.perf_counter() - t0, 1)
                    workloads[name] = res
            except Exception as e:  # a secondary workload must not take the headline line down
                import traceback
                log(traceback.format_exc())
                workloads[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            abx.trim()
            if world > 1:
                dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (bounded sample, measured; N = 1 only) + traversal counters for the byte model ---------
    oracle = oracle_all_threads()
    cpu_baseline = None
    if world == 1:
        cv, cq, cs, _ = cpu_sample_inputs(min(n, args.cpu_n), 1)
        cpu_step(oracle, cv, cq, cs)  # warm-up
        reps = [cpu_step(oracle, cv, cq, cs) for _ in range(3)]
        ct = np.median(np.array([r_[:3] for r_ in reps]), 0)
        cpu_baseline = {"value": 3 * len(cv) / float(ct.sum()) / 1e6, "unit": UNIT, "cores": oracle.num_threads(),
                        "kind": "port", "extrapolated": False,
                        "sample": "the whole step at 1/%d scale, measured: build of %d points + %d radius + %d kNN "
                                  "queries (same density, radius, k; BASELINE configs[0] size), median of 3; "
                                  "oracle/arborx_oracle.cpp (OpenMP)" % (max(1, n // len(cv)), len(cv), len(cv), len(cv)),
                        "build_Mprims_s": len(cv) / ct[0] / 1e6, "radius_Mqueries_s": len(cv) / ct[1] / 1e6,
                        "knn_Mqueries_s": len(cv) / ct[2] / 1e6}
    # counters on the full-size tree of rank 0 (reference-layout tree at the benchmark's depth)
    ctr = cpu_counters(oracle, values, queries, spheres, min(q, args.counter_sample))
    if cpu_baseline is not None:
        cpu_baseline["counters_per_query"] = ctr

    models = kernel_models(ctr, n, q, nnz)
    kernels = kernel_table(prof, models, traffic, peak)
    top = kernels[0] if kernels else None
    roofline = None
    if top and "model_gbs" in top:
        roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["model_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": top["model_frac"], "traffic": top.get("dram_bytes_ncu"), "peak_source": peak_src,
                    "launch_ms": top["max_ms"], "share_of_step": top["total_ms"] / (ms_per_step * args.steps),
                    "achieved_is": "SURVEY 8(d) algorithmic bytes of the full-size launch / its CUDA-event duration "
                                   "(node visits of the reference algorithm, cache hits included)",
                    "dram_achieved": top.get("dram_gbs"), "dram_frac": top.get("dram_frac"),
                    "dram_is": "dram__bytes_read+write of the same kernel from the committed ncu capture "
                               "(profiles/traffic.json) / the live launch duration"}

    # phase figures (SURVEY.md 8(d) / BASELINE.md section 2).  The survey charges the radius phase with two
    # traversal passes (count + fill); this library traverses once and compacts, so both are shown.
    sp_pass = traversal_bytes(ctr["spatial_I"], ctr["spatial_L"])
    b_build = 300.0 * n
    b_radius2 = q * (2 * sp_pass + 16 + 4 + 92) + 4.0 * nnz
    b_radius1 = q * (sp_pass + 16 + 4 + 92) + 4.0 * nnz
    b_knn = q * (traversal_bytes(ctr["nearest_I"], ctr["nearest_L"]) + 16 + 4 + 4 * K_NEIGHBORS + 92)

    def dram_of(names):
        tot = 0.0
        for row in kernels:
            if any(nm in row["kernel"] for nm in names) and "dram_bytes_ncu" in row:
                tot += row["dram_bytes_ncu"] * row["launches"] / args.steps
        return tot or None

    comp = {
        "build_Mprims_s": n / parts[0] / 1e3, "radius_Mqueries_s": q / parts[1] / 1e3,
        "knn_Mqueries_s": q / parts[2] / 1e3,
        "build_ms": parts[0], "radius_ms": parts[1], "knn_ms": parts[2],
        "results_per_radius_query": nnz / q,
        "fractions_note": "model_frac = SURVEY 8(d) bytes / time / peak (reference node-visit traffic, cache hits "
                          "included: not a bound when > 1); dram_frac = ncu DRAM bytes of the phase's kernels / "
                          "time / peak",
    }
    comp.update(phase_fracs("build", b_build, dram_of(["sceneBounds", "morton64", "radixHistogramKernel<u64>",
                                                       "onesweepPassKernel<u64>", "segmentFix", "hierarchy"]),
                            parts[0], peak))
    comp.update(phase_fracs("radius", b_radius2, dram_of(["spatialKernel", "onesweepPassKernel<u32>", "morton32"]),
                            parts[1], peak))
    comp["radius_model_frac_one_pass"] = b_radius1 / (parts[1] * 1e-3) / 1e9 / peak
    comp.update(phase_fracs("knn", b_knn, dram_of(["nearestKernel"]), parts[2], peak))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(n, q, r, world),
        "components": comp,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "api": ("abx_dist_create_host + abx_dist_query_spatial_crs_host + abx_dist_query_nearest_crs_host "
                        "(compact results; numa_pinned=%s)" % numa_pinned) if world > 1 else
                       ("abx_bvh_build_host + abx_query_spatial_crs_host + abx_query_nearest_crs_host (each query "
                        "batch as %d host calls, issued from %d host threads / streams; numa_pinned=%s)"
                        % (args.e2e_chunks, args.e2e_threads, numa_pinned))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels[:12],
        "workloads": workloads,
        "cpu_baseline": cpu_baseline,
        "counters_per_query": ctr,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------ secondary workloads ----
def _time_gpu(torch, fn, warmup, steps):
    # the previous result is dropped before the next call (as a time-stepping caller would): its device blocks go
    # back to the library's cache and the call under test does not pay a driver allocation for a second copy
    out = None
    for _ in range(warmup):
        out = None
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = None
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def _dram_sum(prof, traffic, steps):
    """sum over the kernels of a profiled region of (ncu DRAM bytes per launch) x launches per step; None when a
    kernel of the region has no committed capture (the fraction would be an underestimate)."""
    tot, missing = 0.0, []
    for name, cnt, ms, mx in prof:
        t = traffic.get(name, traffic.get(name.split("<")[0].strip("( ")))
        if t is None:
            if ms / max(sum(p[2] for p in prof), 1e-9) > 0.02:
                missing.append(name)
            continue
        tot += t * cnt / steps
    return (tot if not missing else None), missing


def wl_dbscan(args, abx, torch, dist, space, comm, rank, world, peak, traffic):
    """BASELINE configs[2]: ArborX::dbscan on the GanTao clustered cloud, n = 10M, eps = 200 (N = 1 only; the
    N > 1 form is wl_distributed_dbscan)."""
    if world > 1:
        return None
    from tests import clouds
    n, eps = args.dbscan_n, 200.0
    pts = clouds.gan_tao(3, n)
    d = torch.from_numpy(pts).cuda()
    res = {"config": "GanTao seed spreader n=%d, eps=200, 3-D (benchmarks/cluster/data_timpl.hpp:253-335 restated)" % n}
    steps = max(2, min(args.steps, 5))
    for impl, iname in ((1, "densebox"), (0, "fdbscan")):
        for minpts in (5, 2):
            params = abx.DBSCANParameters(impl, 0)
            labels = abx.dbscan(space, d, eps, minpts, params)  # warm-up
            torch.cuda.synchronize()
            abx.profile_enable(True)
            ms, labels = _time_gpu(torch, lambda: abx.dbscan(space, d, eps, minpts, params), 0, steps)
            prof = abx.profile_report()
            abx.profile_enable(False)
            lab = labels.cpu().numpy()
            kern = {k: round(t / steps, 3) for k, c, t, mx in prof[:6]}
            build_ms = sum(t for k, c, t, mx in prof if any(s in k for s in ("onesweep", "hierarchy", "morton64",
                                                                          "sceneBounds", "radixHistogram",
                                                                          "segmentFix", "radixScan"))) / steps
            dram, missing = _dram_sum(prof, traffic, steps)
            row = {"ms": ms, "Mpoints_s": n / ms / 1e3, "clusters": int(len(np.unique(lab[lab >= 0]))),
                   "noise": int((lab < 0).sum()), "kernels_ms_per_call": kern,
                   "phase_ms": {"construction_kernels": round(build_ms, 3), "query+cluster": round(ms - build_ms, 3)}}
            if dram:
                row["dram_bytes_ncu"] = dram
                row["dram_frac"] = dram / (ms * 1e-3) / 1e9 / peak
            elif missing:
                row["dram_frac"] = None
                row["dram_missing_capture"] = missing[:4]
            res["%s_minpts%d" % (iname, minpts)] = row
    # CPU side: bounded sample (a GanTao cloud of cpu_n points: same local density) + counters for the byte model
    oracle = oracle_all_threads()
    ns = min(n, args.cpu_n)
    sample = clouds.gan_tao(3, ns)
    for impl, iname, minpts in ((1, "densebox", 5), (0, "fdbscan", 5), (0, "fdbscan", 2)):
        t0 = time.perf_counter()
        _, stats = oracle.dbscan(sample, eps, minpts, impl, 0, return_stats=True)
        t_cpu = time.perf_counter() - t0
        key = "%s_minpts%d" % (iname, minpts)
        I, L = float(stats[2]) / ns, float(stats[3]) / ns
        # B_build(20-byte leaves) + sum(32 I + 20 L) over the count and the half traversal + 24 n;
        # DenseBox adds the cell ids / permutation (12 n) and a second 200n-class sort
        model = n * (312.0 + traversal_bytes(I, L) + 24.0 + (212.0 if impl == 1 else 0.0))
        res[key]["model_bytes"] = model
        res[key]["model_frac"] = model / (res[key]["ms"] * 1e-3) / 1e9 / peak
        res[key]["cpu_baseline"] = {"Mpoints_s": ns / t_cpu / 1e6, "cores": oracle.num_threads(), "kind": "port",
                                    "sample": "oracle dbscan on a GanTao cloud of %d points, measured" % ns,
                                    "counters_per_point": {"I": I, "L": L}}
    res["value_Mpoints_s"] = res["densebox_minpts5"]["Mpoints_s"]
    res["reference_sample_note"] = ("the reference README's sample run (HACC 37M, minpts 2, unstated hardware): 161 "
                                    "Mpoints/s, benchmarks/cluster/README.md:87-108 -- different data")
    return res


def wl_bvh_1m(args, abx, torch, dist, space, comm, rank, world, peak, traffic):
    """BASELINE configs[0]: the same step at 1M points / 1M queries -- the size the CPU arm is measured at, so this
    block and `cpu_baseline` are a like-for-like pair.  Single tree: N = 1."""
    if world > 1:
        return None
    n = q = args.cpu_n
    values, queries, spheres, r = make_inputs(n, q)
    dv = torch.from_numpy(values).cuda()
    ps = abx.intersects(torch.from_numpy(spheres).cuda())
    pn = abx.nearest(torch.from_numpy(queries).cuda(), K_NEIGHBORS)
    steps = max(3, args.steps)
    ms_b, bvh = _time_gpu(torch, lambda: abx.BoundingVolumeHierarchy(space, dv), 2, steps)
    ms_r, out = _time_gpu(torch, lambda: bvh.query(space, ps), 2, steps)
    nnz = int(out[0].shape[0])
    ms_k, _ = _time_gpu(torch, lambda: bvh.query(space, pn), 2, steps)
    ms = ms_b + ms_r + ms_k
    return {"config": "build + intersects(sphere) CRS + nearest(k=10) CRS at n = q = %d (same density, radius, k)" % n,
            "ms_per_step": ms, "build_ms": ms_b, "radius_ms": ms_r, "knn_ms": ms_k,
            "value_Mitems_s": 3 * n / ms / 1e3, "build_Mprims_s": n / ms_b / 1e3,
            "radius_Mqueries_s": q / ms_r / 1e3, "knn_Mqueries_s": q / ms_k / 1e3, "results_per_radius_query": nnz / q,
            "note": "the tree (64 MB of nodes) fits the 126 MB L2 at this size; the radius figure includes no lazy "
                    "wide-record conversion (the tree is reused across the timed queries)"}


def wl_mst(args, abx, torch, dist, space, comm, rank, world, peak, traffic):
    """Euclidean minimum spanning tree (k = 1), the mutual-reachability tree HDBSCAN uses
    (k = 5) and hdbscan itself over 10M points: the bench's uniform cloud and the GanTao cluster cloud
    (cluster/ArborX_MinimumSpanningTree.hpp; no published number in BASELINE.md).  Single tree: N = 1."""
    if world > 1:
        return None
    from tests import clouds
    n = args.mst_n
    res = {"config": "MinimumSpanningTree over n=%d points, Boruvka rounds on the BVH; k = 1 Euclidean, k = 5 mutual "
                     "reachability (core distance = 5th nearest, the point itself included)" % n}
    steps = 3
    for cname, make in (("uniform", lambda m: clouds.filled_box(0x5EED0001, m)), ("gantao", lambda m: clouds.gan_tao(3, m))):
        d = torch.from_numpy(np.ascontiguousarray(make(n), np.float32)).cuda()
        for k in (1, 5):
            abx.MinimumSpanningTree(space, d, k)  # first call of the shape: the device-memory cache fills
            torch.cuda.synchronize()
            abx.profile_enable(True)
            ms, mst = _time_gpu(torch, lambda: abx.MinimumSpanningTree(space, d, k), 0, steps)
            prof = abx.profile_report()
            abx.profile_enable(False)
            res["%s_k%d" % (cname, k)] = {
                "ms": ms, "Mpoints_s": n / ms / 1e3, "rounds": mst.iterations,
                "total_weight": float(mst.weights.double().sum().item()),
                "kernels_ms_per_call": {name: round(t / steps, 3) for name, c, t, mx in prof[:5]}}
        # hdbscan(space, points, 5): the hybrid (dendrogram grown with the rounds, on the device) against MST + the
        # union-find pass, whose loop runs on the host in the reference as well
        abx.hdbscan(space, d, 5)
        ms_h, _ = _time_gpu(torch, lambda: abx.hdbscan(space, d, 5), 0, steps)
        ms_u, _ = _time_gpu(torch, lambda: abx.hdbscan(space, d, 5, abx.DENDROGRAM_UNION_FIND), 0, 1)
        res["%s_hdbscan_minpts5" % cname] = {"boruvka_ms": ms_h, "Mpoints_s": n / ms_h / 1e3, "union_find_ms": ms_u}
        del d
    oracle = oracle_all_threads()
    ns = min(n, args.cpu_n)
    for cname, make in (("uniform", lambda m: clouds.filled_box(0x5EED0001, m)), ("gantao", lambda m: clouds.gan_tao(3, m))):
        sample = np.ascontiguousarray(make(ns), np.float32)
        t0 = time.perf_counter()
        oracle.mst(sample, 1)
        t_cpu = time.perf_counter() - t0
        res["%s_k1" % cname]["cpu_baseline"] = {"Mpoints_s": ns / t_cpu / 1e6, "cores": oracle.num_threads(),
                                                "kind": "port",
                                                "sample": "oracle mst (k = 1) of %d points of the same cloud kind, "
                                                          "measured" % ns}
    res["value_Mpoints_s"] = res["gantao_k1"]["Mpoints_s"]
    return res


def wl_triangles(args, abx, torch, dist, space, comm, rank, world, peak, traffic):
    """BASELINE configs[4]: BVH over the 20 971 520 triangles of the icosphere (--refinements 10), nearest(point, 1)
    with distances, intersects(ray) CRS (benchmarks/triangulated_surface_distance/triangulated_surface_distance.cpp
    :182-222, generator.hpp).  Single tree: N = 1."""
    if world > 1:
        return None
    from tests import clouds
    vert, tri = clouds.icosphere(args.tri_refinements)
    soup = clouds.triangle_soup(vert, tri)
    T, V = soup.shape[0], vert.shape[0]
    d_soup = torch.from_numpy(soup).cuda()
    steps = 2
    res = {"config": "unit icosphere, %d refinements: %d triangles (flat 36-byte form), %d vertices"
                     % (args.tri_refinements, T, V)}
    abx.profile_enable(True)
    ms_build, bvh = _time_gpu(torch, lambda: abx.BoundingVolumeHierarchy(space, d_soup, abx.TRIANGLE), 1, steps)
    prof_b = abx.profile_report()
    abx.profile_enable(False)
    res["build"] = {"ms": ms_build, "Mprims_s": T / ms_build / 1e3, "model_bytes": 404.0 * T,
                    "model_frac": 404.0 * T / (ms_build * 1e-3) / 1e9 / peak,
                    "kernels_ms_per_call": {k: round(t / (steps + 1), 3) for k, c, t, mx in prof_b[:6]}}
    # query sets: the benchmark's own (V points uniform in [-cbrt(V), cbrt(V)]^3: far from the surface, where every
    # triangle is nearly equidistant and the search degenerates -- timed on a subset), and points on a shell around
    # the surface (the useful regime), V of them
    far = clouds.filled_box(0x5EED0051, V)[:args.tri_far_queries]
    shell = clouds.shell_points(0x5EED0052, V)
    rays = clouds.ball_rays(0x5EED0053, V)
    oracle = oracle_all_threads()
    t0 = time.perf_counter()
    otree = oracle.Tree(soup, oracle.PRIM_TRI)
    t_cpu_build = time.perf_counter() - t0
    res["build"]["cpu_baseline"] = {"Mprims_s": T / t_cpu_build / 1e6, "cores": oracle.num_threads(), "kind": "port",
                                    "sample": "oracle build of the same %d triangles, measured" % T}
    for name, pts_h, cpu_q in (("nearest_far", far, 20_000), ("nearest_shell", shell, 200_000)):
        dq = torch.from_numpy(pts_h).cuda()
        pred = abx.nearest(dq, 1)
        ms, out = _time_gpu(torch, lambda: bvh.query(space, pred, return_distances=True), 1, steps)
        qn = pts_h.shape[0]
        t0 = time.perf_counter()
        roff, ridx, rd, ctr = otree.nearest_crs(pts_h[:cpu_q], 1, True, counters=True)
        t_cpu = time.perf_counter() - t0
        # parity on the sample: bit-identical distances
        same = bool(np.array_equal(out[2][:cpu_q].cpu().numpy(), rd))
        I, L = float(ctr[0]) / cpu_q, float(ctr[1]) / cpu_q
        model = qn * (traversal_bytes(I, L, 40) + 12 + 4 + 4 + 4 + 92)
        res[name] = {"queries": qn, "ms": ms, "Mqueries_s": qn / ms / 1e3, "model_bytes": model,
                     "model_frac": model / (ms * 1e-3) / 1e9 / peak, "distances_match_oracle_sample": same,
                     "cpu_baseline": {"Mqueries_s": cpu_q / t_cpu / 1e6, "cores": oracle.num_threads(), "kind": "port",
                                      "sample": "oracle nearest(point, 1) on the first %d queries, measured" % cpu_q,
                                      "counters_per_query": {"I": I, "L": L}}}
        del dq, pred, out
    dr = torch.from_numpy(rays).cuda()
    pred = abx.intersects(dr, abx.RAY_PRED)
    ms, out = _time_gpu(torch, lambda: bvh.query(space, pred), 1, steps)
    cpu_q = 200_000
    t0 = time.perf_counter()
    roff, ridx = otree.spatial_crs(rays[:cpu_q], oracle.PRED_RAY, True, 0)
    t_cpu = time.perf_counter() - t0
    _, ctr = otree.spatial_count(rays[:cpu_q], oracle.PRED_RAY, counters=True)
    I, L = float(ctr[0]) / cpu_q, float(ctr[1]) / cpu_q
    nnz = int(out[0].shape[0])
    goff = out[1][:cpu_q + 1].cpu().numpy()
    model = V * (traversal_bytes(I, L, 40) + 24 + 4 + 92) + 4.0 * nnz
    res["ray_intersects"] = {"queries": V, "ms": ms, "Mqueries_s": V / ms / 1e3, "results_per_ray": nnz / V,
                             "model_bytes": model, "model_frac": model / (ms * 1e-3) / 1e9 / peak,
                             "offsets_match_oracle_sample": bool(np.array_equal(goff, roff)),
                             "cpu_baseline": {"Mqueries_s": cpu_q / t_cpu / 1e6, "cores": oracle.num_threads(),
                                              "kind": "port",
                                              "sample": "oracle intersects(ray) CRS on the first %d rays, measured" % cpu_q,
                                              "counters_per_query": {"I": I, "L": L}}}
    return res


def wl_distributed(args, abx, torch, dist, space, comm, rank, world, peak, traffic):
    """BASELINE configs[3] at its stated size: DistributedTree over 100M points in total, sharded over the N GPUs on
    the touching block lattice of distributed_tree_driver (strong scaling: n_r = 100M / N values and as many queries
    per rank; queries from the same distribution; r = cbrt(10 * 6 / pi), k = 10)."""
    from arborx_b200.distributed import Communicator, DistributedTree
    from tests import clouds
    total = args.dist_total
    n_r = total // world
    if comm is None:
        comm = Communicator.local_group(1)[0]
    a = float(np.float32(np.cbrt(float(n_r))))
    off = torch.from_numpy(block_offset(n_r, rank, world)).cuda()
    vals = clouds.filled_box_torch(0x5EED0101 + 16 * rank, n_r, "cuda", a) + off
    qs = clouds.filled_box_torch(0x5EED0102 + 16 * rank, n_r, "cuda", a) + off
    r = float(clouds.bvh_driver_radius(K_NEIGHBORS))
    spheres = torch.cat([qs, torch.full((n_r, 1), r, device="cuda")], 1).contiguous()
    steps = max(2, min(args.steps, 3))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    times = []
    nnz = 0
    for it in range(1 + steps):
        e = [ev() for _ in range(4)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e[0].record()
        tree = DistributedTree(comm, space, vals)
        e[1].record()
        v, o = tree.query(space, abx.intersects(spheres))
        e[2].record()
        nnz = v.shape[0]
        del v, o
        kv, ko = tree.query(space, abx.nearest(qs, K_NEIGHBORS))
        e[3].record()
        torch.cuda.synchronize()
        del kv, ko, tree
        if it > 0:
            times.append([e[i].elapsed_time(e[i + 1]) for i in range(3)])
    t = torch.tensor(np.array(times).mean(0), device="cuda", dtype=torch.float64)
    cnt = torch.tensor([float(nnz)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    tb, tr, tk = [float(x) for x in t.tolist()]
    step_ms = tb + tr + tk
    return {"config": "%d points + %d radius + %d kNN queries in total, %d per rank on %d GPU(s), touching blocks"
                      % (n_r * world, n_r * world, n_r * world, n_r, world),
            "scaling": "strong", "ms_per_step": step_ms, "build_ms": tb, "radius_ms": tr, "knn_ms": tk,
            "value_Mitems_s": 3.0 * n_r * world / step_ms / 1e3,
            "build_Mprims_s": n_r * world / tb / 1e3, "radius_Mqueries_s": n_r * world / tr / 1e3,
            "knn_Mqueries_s": n_r * world / tk / 1e3, "results_per_radius_query": float(cnt.item()) / (n_r * world),
            "build_model_frac": 300.0 * n_r / (tb * 1e-3) / 1e9 / peak}


def wl_distributed_dbscan(args, abx, torch, dist, space, comm, rank, world, peak, traffic):
    """BASELINE configs[2] at N > 1: distributed DBSCAN (halo exchange + label merge), the GanTao cloud of n points on
    every rank shifted into its slab (touching slabs), weak scaling."""
    if world == 1:
        return None
    from arborx_b200.distributed_dbscan import dbscan as dist_dbscan
    from tests import clouds
    n, eps = args.dbscan_n, 200.0
    pts = clouds.gan_tao(3, n)
    pts[:, 0] += np.float32(rank * 1.0e6)
    d = torch.from_numpy(pts).cuda()
    res = {"config": "GanTao n=%d per rank in touching slabs, eps=200; scaling weak" % n}
    steps = max(2, min(args.steps, 3))
    for impl, iname, minpts in ((1, "densebox", 5), (0, "fdbscan", 2)):
        params = abx.DBSCANParameters(impl, 0)
        labels = dist_dbscan(dist.group.WORLD, space, d, eps, minpts, params)
        dist.barrier()
        ms, labels = _time_gpu(torch, lambda: dist_dbscan(dist.group.WORLD, space, d, eps, minpts, params), 0, steps)
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        res["%s_minpts%d" % (iname, minpts)] = {"ms": ms, "Mpoints_s": world * n / ms / 1e3}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--q", type=int, default=None)
    ap.add_argument("--cpu-n", type=int, default=1_000_000, help="points (= queries) of the CPU arm's per-step sample")
    ap.add_argument("--counter-sample", type=int, default=200_000,
                    help="queries the oracle counts node visits on (byte model)")
    ap.add_argument("--mst-n", type=int, default=10_000_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-threads", type=int, default=3, help="host threads (streams) of the end-to-end path")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="host calls per query batch in the end-to-end path")
    ap.add_argument("--skip-workloads", action="store_true", help="headline step only")
    ap.add_argument("--dbscan-n", type=int, default=10_000_000)
    ap.add_argument("--tri-refinements", type=int, default=10)
    ap.add_argument("--tri-far-queries", type=int, default=262_144)
    ap.add_argument("--dist-total", type=int, default=100_000_000)
    args = ap.parse_args()
    if args.q is None:
        args.q = args.n
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
