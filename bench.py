#!/usr/bin/env python
"""bench.py -- the bvh_driver workload of BASELINE.json (configs[1]) on B200.

One "step" = one pass of the hot path over one batch of synthetic input:
    BVH build over n points  +  intersects(sphere) CRS query with q spheres
    +  nearest(k=10) CRS query with q points
with n = q = 10M filled-box points, r = cbrt(10*6/pi), predicates Morton-sorted,
buffer_size = 0 (benchmarks/bvh_driver/benchmark_registration.hpp:129-205).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

prints ONE JSON line (see the task contract).  `value` is the aggregate item rate
(n + 2q items per step) with inputs resident in HBM; `components` holds the three
rates BASELINE.json names (build Mprims/s, radius Mqueries/s, kNN Mqueries/s) and
their HBM-roofline fractions; `e2e` is the same step through the host-buffer C-ABI
entry points (pinned host inputs, results copied back to the host).

--impl reference times the CPU restatement of the reference (oracle/, OpenMP, all
host threads) -- the real reference cannot be built here (Kokkos absent).
Multi-GPU (--gpus N under torchrun): the path shards in DistributedTree -- every rank owns
n points / q queries of a touching block lattice (distributed_tree_driver layout), the step
is DistributedTree build + distributed radius + distributed kNN, weak scaling, NCCL
all-to-all-v for forwarded queries and results.
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
# stdout must carry exactly one JSON line: keep NCCL's version banner out of it

# stdout must carry exactly ONE JSON line.  Libraries write banners to fd 1 (NCCL prints its
# version there), so fd 1 is pointed at stderr for the whole run and the result line is written
# to the saved descriptor by emit().
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "BVH build Mprims/s; radius & kNN(k=10) Mqueries/s at 10M pts; % HBM roofline"
UNIT = "Mitems/s (n prims + q radius queries + q kNN queries per step second)"
K_NEIGHBORS = 10


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


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


def make_inputs(n, q, rank=None, world=1):
    """rank is None: the single-tree bvh_driver clouds.  Otherwise this rank's block of the
    distributed_tree_driver layout (benchmarks/distributed_tree_driver/distributed_tree_driver.cpp:44-149):
    blocks of side 2a, a = cbrt(n), on a ceil(cbrt(R))^3 lattice, shift = 1 (touching)."""
    from tests import clouds
    if rank is None:
        values = clouds.filled_box(0x5EED0001, n)
        queries = clouds.filled_box(0x5EED0002, q)
    else:
        nb = int(np.ceil(np.cbrt(world) - 1e-9))
        ijk = np.array([rank % nb, (rank // nb) % nb, rank // (nb * nb)], np.float32)
        a = np.float32(np.cbrt(float(n)))
        off = (np.float32(2) * a * ijk).astype(np.float32)
        values = (clouds.filled_box(0x5EED0001 + 16 * rank, n) + off).astype(np.float32)
        queries = (clouds.filled_box(0x5EED0002 + 16 * rank, q) + off).astype(np.float32)
    r = clouds.bvh_driver_radius(K_NEIGHBORS)
    spheres = np.concatenate([queries, np.full((q, 1), r, np.float32)], 1).astype(np.float32)
    return values, queries, spheres, float(r)


# ------------------------------------------------------------------ CPU arm ----
def cpu_run(values, queries, spheres, q_sample, build_n=None):
    """Times the oracle (restated reference, OpenMP) on a bounded sample: full build,
    q_sample of the queries.  Returns rates, the per-query traversal counters used for
    the algorithmic-byte figures, and the wall time."""
    import oracle
    n = len(values) if build_n is None else build_n
    t0 = time.time()
    tree = oracle.Tree(values[:n])
    t_build = time.time() - t0
    sp = spheres[:q_sample]
    t0 = time.time()
    off, idx = tree.spatial_crs(sp, oracle.PRED_SPHERE, True, 0)
    t_radius = time.time() - t0
    t0 = time.time()
    koff, kidx, kd = tree.nearest_crs(queries[:q_sample], K_NEIGHBORS, True)
    t_knn = time.time() - t0
    _, c_sp = tree.spatial_count(sp, counters=True)
    _, _, _, c_nn = tree.nearest_crs(queries[:q_sample], K_NEIGHBORS, True, counters=True)
    return dict(n=n, q_sample=q_sample, t_build=t_build, t_radius=t_radius, t_knn=t_knn,
                nnz_per_query=float(off[-1]) / q_sample,
                spatial_I=float(c_sp[0]) / q_sample, spatial_L=float(c_sp[1]) / q_sample,
                nearest_I=float(c_nn[0]) / q_sample, nearest_L=float(c_nn[1]) / q_sample,
                cores=oracle.num_threads())


def combined_rate(n, q, t_build, t_radius, t_knn):
    return (n + 2 * q) / (t_build + t_radius + t_knn) / 1e6


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, q = args.n, args.q
    values, queries, spheres, r = make_inputs(n, q)
    qs = min(q, args.cpu_sample)
    times = []
    res = None
    for it in range(args.warmup + args.steps):
        res = cpu_run(values, queries, spheres, qs)
        scale = q / qs
        step = res["t_build"] + (res["t_radius"] + res["t_knn"]) * scale
        if it >= args.warmup:
            times.append((step, res["t_build"], res["t_radius"] * scale, res["t_knn"] * scale))
    t = np.mean(np.array(times), 0)
    value = (n + 2 * q) / t[0] / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t[0] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n, q, r),
        "components": {"build_Mprims_s": n / t[1] / 1e6, "radius_Mqueries_s": q / t[2] / 1e6,
                       "knn_Mqueries_s": q / t[3] / 1e6},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port",
                         "sample": "per step: full build of n=%d points, radius+kNN on the first %d of %d queries "
                                   "(query time scaled by q/sample); oracle/arborx_oracle.cpp, OpenMP" % (n, qs, q)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n, q, r):
    return {"workload": "bvh_driver filled_box: build + intersects(sphere) CRS + nearest(k=10) CRS "
                        "(BASELINE.json configs[1])",
            "n_values": n, "n_queries": q, "k": K_NEIGHBORS, "radius": r, "sort_predicates": True, "buffer_size": 0,
            "cloud": "uniform in [-cbrt(n), cbrt(n)]^3, counter-based RNG (tests/clouds.py)",
            "cache": "inputs and tree (%.0f MB nodes) larger than the 126 MB L2; no explicit flush" % (64e-6 * n)}


# ------------------------------------------------------------------ GPU arm ----
def algorithmic_bytes(cpu, n, q, nnz):
    """Per-launch algorithmic bytes of each kernel:
    traversals are defined on the reference-layout tree from the oracle's counters
    (32 B per internal-node test, 20 B per leaf test)."""
    sp = 32.0 * cpu["spatial_I"] + 20.0 * cpu["spatial_L"]
    nn = 32.0 * cpu["nearest_I"] + 20.0 * cpu["nearest_L"]
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
        "hierarchyKernel": 64.0 * n,
        "radixHistogramKernel<u64>": 8.0 * n,
        "morton64Kernel": 20.0 * n,
        "sceneBoundsKernel": 12.0 * n,
    }


def run_ours(args):
    import torch
    import torch.distributed as dist

    import arborx_b200 as abx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
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

    if world > 1:
        from arborx_b200.distributed import DistributedTree
        comm = dist.group.WORLD

    def make_tree(vals):
        # N > 1: the path shards in DistributedTree (primitives per GPU, replicated top tree,
        # all-to-all-v of forwarded queries and results); N = 1: the single tree
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
        print("per-step (build, radius, knn) ms:\n", per_step, file=sys.stderr)
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
    hp_spatial = abx.intersects(h_spheres)
    hp_nearest = abx.nearest(h_queries, K_NEIGHBORS)

    pinned_out = {}
    import concurrent.futures
    import threading
    E2E_THREADS, E2E_CHUNKS = args.e2e_threads, args.e2e_chunks
    e2e_pool = concurrent.futures.ThreadPoolExecutor(E2E_THREADS)
    e2e_copy_stream = torch.cuda.Stream()
    e2e_local = threading.local()
    bounds = [q * c // E2E_CHUNKS for c in range(E2E_CHUNKS + 1)]
    hp_spatial_parts = [abx.intersects(h_spheres[bounds[c]:bounds[c + 1]]) for c in range(E2E_CHUNKS)]
    hp_nearest_parts = [abx.nearest(h_queries[bounds[c]:bounds[c + 1]], K_NEIGHBORS) for c in range(E2E_CHUNKS)]

    def e2e_task(bvh, preds):
        if not hasattr(e2e_local, "space"):
            torch.cuda.set_device(local_rank)
            e2e_local.space = abx.ExecutionSpace(torch.cuda.Stream())
        idx, off = bvh.query(e2e_local.space, preds)
        return int(off[-1]), idx.numel()

    def e2e_step():
        if world > 1:
            # DistributedTree takes device data: the host<->device copies are done here, inside the step
            # all three uploads are queued up front; the radius results go back on a copy stream while the kNN
            # query runs
            dv = h_values.cuda(non_blocking=True)
            d_sp = h_spheres.cuda(non_blocking=True)
            d_qq = h_queries.cuda(non_blocking=True)
            tree = make_tree(dv)
            main = torch.cuda.current_stream()

            def to_host(pairs, stream):
                hs = []
                with torch.cuda.stream(stream):
                    for name, t in pairs:
                        buf = pinned_out.get(name)
                        if buf is None or buf.numel() < t.numel():
                            buf = torch.empty(int(t.numel() * 1.1) + 16, dtype=t.dtype, pin_memory=True)
                            pinned_out[name] = buf
                        h = buf[:t.numel()].view(t.shape)
                        h.copy_(t, non_blocking=True)
                        t.record_stream(stream)
                        hs.append(h)
                return hs

            idx, off = tree.query(space, abx.intersects(d_sp))
            e2e_copy_stream.wait_stream(main)
            outs = to_host((("idx", idx), ("off", off)), e2e_copy_stream)
            kidx, koff = tree.query(space, abx.nearest(d_qq, K_NEIGHBORS))
            outs += to_host((("kidx", kidx), ("koff", koff)), main)
            torch.cuda.synchronize()
            idx, off, kidx, koff = outs
            return int(off[-1]) + int(koff[-1]), idx.numel(), kidx.numel()
        bvh = abx.BoundingVolumeHierarchy(space, h_values)
        space.fence()
        # The query batches are independent, and so are their parts: each batch is issued as E2E_CHUNKS host
        # calls from E2E_THREADS host threads, every thread on its own execution space instance (stream), so
        # the result copy of one call overlaps the traversal of the next (PCIe is the long pole of this path:
        # 0.4 GB in, 0.88 GB out per step).  kNN parts first: they are the longer ones.
        futures = [e2e_pool.submit(e2e_task, bvh, p) for p in hp_nearest_parts + hp_spatial_parts]
        total, n_idx, n_kidx = 0, 0, 0
        for f, p in zip(futures, hp_nearest_parts + hp_spatial_parts):
            last, cnt = f.result()
            total += last
            if p.tag == "nearest":
                n_kidx += cnt
            else:
                n_idx += cnt
        # results are host tensors: their last offsets were read on the host inside the tasks
        return total, n_idx, n_kidx

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, n_idx, n_kidx = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * (n + 2 * q) / e2e_s / 1e6
    h2d = 12 * n + 16 * q + 12 * q
    d2h = 4 * (q + 1) + 4 * n_idx + 4 * (q + 1) + 4 * n_kidx

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (bounded sample) + traversal counters for the roofline ---------
    qs = min(q, args.cpu_sample if world == 1 else min(args.cpu_sample, 100_000))
    cpu = cpu_run(values, queries, spheres, qs)  # rank 0's block when N > 1 (counters for the byte model)
    scale = q / qs
    cpu_value = combined_rate(n, q, cpu["t_build"], cpu["t_radius"] * scale, cpu["t_knn"] * scale)

    peak, peak_src, _ = peaks()
    alg = algorithmic_bytes(cpu, n, q, nnz)
    # kernel table from the live CUDA-event profile of the timed region
    kernels = []
    for name, cnt, ms in prof:
        key = None
        for kname in alg:  # tags are resolved names, e.g. "spatialKernel<count>", "(hierarchyKernel<K>)"
            if kname in name or ("<" not in kname and kname in name.split("<")[0]):
                key = kname
        per_launch_ms = ms / cnt
        row = {"kernel": name, "launches": cnt, "total_ms": round(ms, 4), "avg_ms": round(per_launch_ms, 5)}
        if key:
            row["algorithmic_bytes"] = alg[key]
            row["achieved_gbs"] = alg[key] / (per_launch_ms * 1e-3) / 1e9
        kernels.append(row)
    top = kernels[0] if kernels else None
    roofline = None
    if top and "achieved_gbs" in top:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(top["kernel"].split("<")[0].strip("( "))
        roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": top["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                    "share_of_step": top["total_ms"] / (ms_per_step * args.steps)}

    # phase rooflines (SURVEY.md 8(d) / BASELINE.md section 2)
    b_build = 300.0 * n
    sp_pass = 32.0 * cpu["spatial_I"] + 20.0 * cpu["spatial_L"]
    b_radius = q * (2 * sp_pass + 16 + 4 + 92) + 4.0 * nnz
    b_knn = q * (32.0 * cpu["nearest_I"] + 20.0 * cpu["nearest_L"] + 16 + 4 + 4 * K_NEIGHBORS + 92)
    comp = {
        "build_Mprims_s": n / parts[0] / 1e3, "radius_Mqueries_s": q / parts[1] / 1e3,
        "knn_Mqueries_s": q / parts[2] / 1e3,
        "build_ms": parts[0], "radius_ms": parts[1], "knn_ms": parts[2],
        "build_roofline_frac": b_build / (parts[0] * 1e-3) / 1e9 / peak,
        "radius_roofline_frac": b_radius / (parts[1] * 1e-3) / 1e9 / peak,
        "knn_roofline_frac": b_knn / (parts[2] * 1e-3) / 1e9 / peak,
        "algorithmic_bytes": {"build": b_build, "radius": b_radius, "knn": b_knn},
        "results_per_radius_query": nnz / q,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(n, q, r),
        "components": comp,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                "api": "abx_bvh_build_host + abx_query_spatial_crs_host + abx_query_nearest_crs_host (each query batch as %d host calls, issued from %d host threads / streams)" % (args.e2e_chunks, args.e2e_threads)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels[:12],
        "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                         "sample": "full build of n=%d points; radius+kNN on the first %d of %d queries, query time "
                                   "scaled by q/sample; oracle/arborx_oracle.cpp (OpenMP)" % (n, qs, q),
                         "build_Mprims_s": n / cpu["t_build"] / 1e6,
                         "radius_Mqueries_s": qs / cpu["t_radius"] / 1e6, "knn_Mqueries_s": qs / cpu["t_knn"] / 1e6,
                         "counters_per_query": {k: cpu[k] for k in ("spatial_I", "spatial_L", "nearest_I",
                                                                    "nearest_L", "nnz_per_query")}},
    }
    if world > 1:
        line["config"]["parallelism"] = ("DistributedTree over %d GPUs: %d points and %d queries per rank on a touching "
                                         "block lattice; top tree + all-to-all-v forwarding (NCCL)" % (world, n, q))
        line["config"]["workload"] += " through DistributedTree (BASELINE.json configs[3] layout, weak scaling)"
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_dbscan_distributed(args, world, rank, local_rank):
    """`--workload dbscan --gpus N` (N > 1): the distributed DBSCAN (halo exchange + label merge) with the
    GanTao cloud of n points on every rank, rank r's copy shifted by r * L along x (touching slabs), weak scaling."""
    import torch
    import torch.distributed as dist

    import arborx_b200 as abx
    from arborx_b200.distributed_dbscan import dbscan as dist_dbscan
    from tests import clouds
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, eps = args.n, 200.0
    # the same cloud on every rank (identical work per GPU: the step time then shows the halo / merge
    # overhead, not the luck of a seed), shifted into the rank's slab
    pts = clouds.gan_tao(3, n)
    pts[:, 0] += np.float32(rank * 1.0e6)
    space = abx.ExecutionSpace()
    d = torch.from_numpy(pts).cuda()
    res = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for impl, iname, minpts in ((1, "densebox", 5), (0, "fdbscan", 2)):
        params = abx.DBSCANParameters(impl, 0)
        for _ in range(max(1, args.warmup)):
            labels = dist_dbscan(dist.group.WORLD, space, d, eps, minpts, params)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(args.steps):
            labels = dist_dbscan(dist.group.WORLD, space, d, eps, minpts, params)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        nclu = torch.unique(labels[labels >= 0])
        sizes = [None] * world
        dist.all_gather_object(sizes, nclu.cpu().numpy())
        res["%s_minpts%d" % (iname, minpts)] = {"ms": ms, "Mpoints_s": world * n / ms / 1e3,
                                               "clusters": int(len(np.unique(np.concatenate(sizes))))}
    if rank == 0:
        k = "densebox_minpts5"
        emit({"metric": "DBSCAN Mpoints/s (GanTao clustered 3-D, eps=200)", "value": res[k]["Mpoints_s"],
              "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": res[k]["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "f32", "data": "synthetic",
              "config": {"workload": "distributed dbscan, GanTao n=%d per rank in touching slabs, eps=200 "
                                     "(BASELINE.json configs[2] at N GPUs)" % n}, "components": res})
    dist.destroy_process_group()


def run_dbscan(args):
    """Secondary workload (BASELINE.json configs[2]): ArborX::dbscan on a GanTao clustered cloud,
    eps = 200, minpts in {2, 5}, FDBSCAN and FDBSCAN-DenseBox.  One JSON line; `value` = points/s of
    the reference's default configuration (FDBSCAN-DenseBox, minpts = 5)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return run_dbscan_distributed(args, world, int(os.environ.get("RANK", "0")),
                                      int(os.environ.get("LOCAL_RANK", "0")))
    import torch

    import arborx_b200 as abx
    import oracle
    from tests import clouds
    n = args.n
    eps = 200.0
    pts = clouds.gan_tao(3, n)
    space = abx.ExecutionSpace()
    d = torch.from_numpy(pts).cuda()
    res = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for impl, iname in ((1, "densebox"), (0, "fdbscan")):
        for minpts in (5, 2):
            params = abx.DBSCANParameters(impl, 0)
            for _ in range(max(1, args.warmup)):
                labels = abx.dbscan(space, d, eps, minpts, params)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            abx.profile_enable(True)
            e0.record()
            for _ in range(args.steps):
                labels = abx.dbscan(space, d, eps, minpts, params)
            e1.record()
            torch.cuda.synchronize()
            prof = abx.profile_report()
            abx.profile_enable(False)
            ms = e0.elapsed_time(e1) / args.steps
            lab = labels.cpu().numpy()
            res["%s_minpts%d" % (iname, minpts)] = {"ms": ms, "Mpoints_s": n / ms / 1e3,
                                                   "clusters": int(len(np.unique(lab[lab >= 0]))),
                                                   "noise": int((lab < 0).sum()),
                                                   "kernels_ms_per_call": {k: round(t / args.steps, 3)
                                                                           for k, c, t in prof[:8]}}
    ns = min(n, args.cpu_sample * 2)
    t0 = time.time()
    ref = oracle.dbscan(pts[:ns], eps, 5, 1, 0)
    t_cpu = time.time() - t0
    line = {"metric": "DBSCAN Mpoints/s (GanTao clustered 3-D, eps=200)", "value": res["densebox_minpts5"]["Mpoints_s"],
            "unit": "Mpoints/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["densebox_minpts5"]["ms"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "dbscan GanTao n=%d eps=200 (BASELINE.json configs[2])" % n}, "components": res,
            "cpu_baseline": {"value": ns / t_cpu / 1e6, "unit": "Mpoints/s", "cores": oracle.num_threads(),
                             "kind": "port", "sample": "oracle FDBSCAN-DenseBox minpts=5 on the first %d points" % ns}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--q", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=500_000, help="queries timed on the CPU baseline")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-threads", type=int, default=3, help="host threads (streams) of the end-to-end path")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="host calls per query batch in the end-to-end path")
    ap.add_argument("--workload", default="bvh", choices=["bvh", "dbscan"],
                    help="bvh: the headline bvh_driver step (default); dbscan: secondary DBSCAN workload")
    args = ap.parse_args()
    if args.q is None:
        args.q = args.n
    if args.workload == "dbscan":
        run_dbscan(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
