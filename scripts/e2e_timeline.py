#!/usr/bin/env python
"""Where does the end-to-end step of bench.py spend its time?  Host-side timeline of the host-buffer calls
(abx_bvh_build_host, abx_query_spatial_crs_host, abx_query_nearest_crs_host) of one 10M/10M step: every task's
start/end on the host clock, and the step time of partial steps (build only, kNN parts only, spatial parts only),
for several (chunks, threads) splits.  Run on the GPU box; writes gpurun_out/r02_e2e_timeline.log."""
import concurrent.futures
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import arborx_b200 as abx
import bench

K = bench.K_NEIGHBORS


def main():
    n = q = int(os.environ.get("E2E_N", "10000000"))
    values, queries, spheres, r = bench.make_inputs(n, q, None, 1)
    torch.cuda.set_device(0)
    space = abx.ExecutionSpace()
    h_values = torch.from_numpy(values).pin_memory()
    h_spheres = torch.from_numpy(spheres).pin_memory()
    h_queries = torch.from_numpy(queries).pin_memory()
    out = open("gpurun_out/r02_e2e_timeline.log", "w")

    def log(*a):
        s = " ".join(str(x) for x in a)
        print(s)
        out.write(s + "\n")
        out.flush()

    def run(chunks, threads, what, show=False, order="knn_first"):
        pool = concurrent.futures.ThreadPoolExecutor(threads)
        local = threading.local()
        b = [q * c // chunks for c in range(chunks + 1)]
        sp = [abx.intersects(h_spheres[b[c]:b[c + 1]]) for c in range(chunks)]
        nn = [abx.nearest(h_queries[b[c]:b[c + 1]], K) for c in range(chunks)]
        pools = [abx.HostBufferPool() for _ in range(2 * chunks)]
        t_base = [0.0]
        marks = []

        def task(bvh, preds, slot):
            if not hasattr(local, "space"):
                torch.cuda.set_device(0)
                local.space = abx.ExecutionSpace(torch.cuda.Stream())
            t0 = time.perf_counter()
            idx, off = bvh.query(local.space, preds, out=pools[slot])
            t1 = time.perf_counter()
            marks.append((preds.tag, slot, threading.get_ident() % 1000, (t0 - t_base[0]) * 1e3, (t1 - t_base[0]) * 1e3))
            return int(off[-1])

        def step():
            marks.clear()
            t_base[0] = time.perf_counter()
            bvh = abx.BoundingVolumeHierarchy(space, h_values)
            space.fence()
            t_build = (time.perf_counter() - t_base[0]) * 1e3
            parts = []
            if "k" in what:
                parts += nn
            if "s" in what:
                parts += sp
            if order == "interleave" and what == "ks":
                parts = [p for pair in zip(nn, sp) for p in pair]
            if order == "spatial_first" and what == "ks":
                parts = sp + nn
            fs = [pool.submit(task, bvh, p, i) for i, p in enumerate(parts)]
            for f in fs:
                f.result()
            return t_build

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            tb = step()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        log("chunks=%d threads=%d what=%s order=%s: step %.2f ms (min %.2f)  build+fence %.2f" %
            (chunks, threads, what, order, float(np.median(ts)), min(ts), tb))
        if show:
            for m in sorted(marks, key=lambda m: m[3]):
                log("    %-8s slot %2d thr %3d  %7.2f -> %7.2f  (%.2f ms)" % (m[0], m[1], m[2], m[3], m[4], m[4] - m[3]))
        pool.shutdown()

    run(4, 3, "", show=False)
    run(1, 1, "k", show=True)
    run(1, 1, "s", show=True)
    run(4, 3, "k", show=True)
    run(4, 3, "s", show=True)
    run(4, 3, "ks", show=True)
    run(4, 2, "ks", show=True)
    run(4, 3, "ks", order="interleave", show=True)
    run(4, 3, "ks", order="spatial_first")
    run(2, 2, "ks")
    run(8, 2, "ks")
    run(8, 4, "ks", order="interleave")
    run(16, 3, "ks", order="interleave")


if __name__ == "__main__":
    main()
