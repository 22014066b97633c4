"""Staged bring-up on the GPU box: checks every stage against the oracle and prints
where the first divergence is, then rough timings at 10M.  Not a test, not a bench."""
import ctypes as C
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
import oracle  # noqa: E402
from arborx_b200 import _lib  # noqa: E402
from tests import clouds  # noqa: E402
from tests.cuda_engine import CudaEngine  # noqa: E402

L = _lib.lib()
eng = CudaEngine()
space = eng.space


def stage(name, fn):
    t = time.time()
    try:
        fn()
        torch.cuda.synchronize()
        print("[ OK ] %-40s %.2fs" % (name, time.time() - t), flush=True)
    except Exception:
        print("[FAIL] %-40s" % name, flush=True)
        traceback.print_exc()
        sys.stdout.flush()


def first_diff(a, b, what):
    if a.shape != b.shape:
        raise AssertionError("%s: shape %s vs %s" % (what, a.shape, b.shape))
    d = np.nonzero(a != b)[0]
    if len(d):
        i = d[0]
        raise AssertionError("%s: %d mismatches, first at %d: got %s want %s" % (what, len(d), i, a[i], b[i]))


def t_sort():
    for n in (1, 100, 4096, 4097, 50_000, 1_000_003):
        rng = np.random.default_rng(n)
        keys = rng.integers(0, 2 ** 63, n, dtype=np.uint64)
        k, p = eng.sort_u64(keys)
        o = np.argsort(keys, kind="stable")
        first_diff(k, keys[o], "sort_u64 keys n=%d" % n)
        first_diff(p, o.astype(np.uint32), "sort_u64 perm n=%d" % n)
        k32 = rng.integers(0, 2 ** 30, n, dtype=np.uint32)
        k, p = eng.sort_u32(k32)
        o = np.argsort(k32, kind="stable")
        first_diff(k, k32[o], "sort_u32 keys n=%d" % n)
        first_diff(p, o.astype(np.uint32), "sort_u32 perm n=%d" % n)


pts = clouds.filled_box(1, 200_000)


def t_bounds():
    first_diff(eng.scene_bounds(pts), oracle.scene_bounds(pts), "bounds")


def t_codes():
    b = oracle.scene_bounds(pts)
    first_diff(eng.morton64_codes(pts, b), oracle.morton64_codes(pts, b), "codes")


def t_tree():
    for n in (2, 3, 17, 1000, 200_000):
        tc, to = eng.build(pts[:n]), oracle.Tree(pts[:n])
        dc, do = tc.export(), to.export()
        for key in ("codes", "leaf_index", "leaf_rope", "left_child", "rope"):
            first_diff(dc[key], do[key], "tree n=%d %s" % (n, key))
        first_diff(dc["boxes"].ravel(), do["boxes"].ravel(), "tree n=%d boxes" % n)


def t_spatial():
    tc, to = eng.build(pts), oracle.Tree(pts)
    q = clouds.filled_box(2, 50_000)
    sp = np.concatenate([q, np.full((len(q), 1), clouds.bvh_driver_radius(10), np.float32)], 1)
    cnt = tc.spatial_count(sp)
    first_diff(cnt, to.spatial_count(sp), "spatial counts")
    oc, ic = tc.spatial_crs(sp)
    oo, io = to.spatial_crs(sp)
    first_diff(oc, oo, "spatial offsets")
    row = np.repeat(np.arange(len(q)), np.diff(oo))
    first_diff(ic[np.lexsort((ic, row))], io[np.lexsort((io, row))], "spatial indices")


def t_nearest():
    tc, to = eng.build(pts), oracle.Tree(pts)
    q = clouds.filled_box(2, 20_000)
    for k in (1, 10, 40):
        oc, ic, dc = tc.nearest_crs(q, k)
        oo, io, do = to.nearest_crs(q, k)
        first_diff(oc, oo, "knn offsets k=%d" % k)
        first_diff(dc, do, "knn distances k=%d" % k)
        print("   k=%d index mismatches (ties allowed): %d" % (k, int((ic != io).sum())))


def t_half():
    tc, to = eng.build(pts[:30000]), oracle.Tree(pts[:30000])
    pc, po = tc.half_pairs(2.0), to.half_pairs(2.0)
    canon = lambda p: np.unique(np.sort(p.astype(np.int64), 1), axis=0)
    assert len(pc) == len(po), (len(pc), len(po))
    first_diff(canon(pc).ravel(), canon(po).ravel(), "half pairs")


def t_dbscan():
    x = clouds.clustered(11, 100_000, n_clusters=8, domain=1.0e4, spread=20.0, noise_frac=0.02)
    for impl in (0, 1):
        for minpts in (2, 5):
            lab = eng.dbscan(x, 25.0, minpts, impl, 0)
            ref, core = oracle.dbscan(x, 25.0, minpts, 0, 0, return_core=True)
            first_diff(lab[core], ref[core], "dbscan impl=%d minpts=%d core labels" % (impl, minpts))
            first_diff((lab == -1).astype(np.int32), (ref == -1).astype(np.int32), "dbscan noise")


def t_perf():
    n = 10_000_000
    x = torch.from_numpy(clouds.filled_box(0x5EED0001, n)).cuda()
    qv = torch.from_numpy(clouds.filled_box(0x5EED0002, n)).cuda()
    r = float(clouds.bvh_driver_radius(10))
    sp = torch.cat([qv, torch.full((n, 1), r, device="cuda")], 1).contiguous()

    def timed(name, fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        best = 1e9
        for _ in range(reps):
            ev[0].record()
            out = fn()
            ev[1].record()
            torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]))
        print("   %-28s %8.3f ms  (%.1f M/s)" % (name, best, n / best / 1e3), flush=True)
        return out

    b = torch.empty(6, device="cuda")
    codes = torch.empty(n, dtype=torch.int64, device="cuda")
    perm = torch.empty(n, dtype=torch.int32, device="cuda")
    timed("scene_bounds", lambda: _lib.check(L.abx_scene_bounds(space.handle, 0, C.c_void_p(x.data_ptr()), n, C.c_void_p(b.data_ptr()))))
    timed("morton64", lambda: _lib.check(L.abx_morton64(space.handle, 0, C.c_void_p(x.data_ptr()), n, C.c_void_p(b.data_ptr()), C.c_void_p(codes.data_ptr()))))
    c2 = codes.clone()
    timed("sort_u64 (sorted input after 1st)", lambda: _lib.check(L.abx_sort_u64(space.handle, C.c_void_p(c2.data_ptr()), C.c_void_p(perm.data_ptr()), n)))

    def sort_fresh():
        c3 = codes.clone()
        _lib.check(L.abx_sort_u64(space.handle, C.c_void_p(c3.data_ptr()), C.c_void_p(perm.data_ptr()), n))
    timed("clone + sort_u64 (random)", sort_fresh)
    timed("clone only", lambda: codes.clone())
    bvh = timed("build (whole)", lambda: abx.BoundingVolumeHierarchy(space, x))
    timed("spatial count (sorted preds)", lambda: bvh.count(space, abx.intersects(sp)))
    res = timed("spatial CRS", lambda: bvh.query(space, abx.intersects(sp)))
    print("   nnz =", res[0].numel())
    timed("kNN k=10", lambda: bvh.query(space, abx.nearest(qv, 10)))
    timed("kNN k=10 unsorted", lambda: bvh.query(space, abx.nearest(qv, 10), abx.TraversalPolicy(0, False)))
    xc = torch.from_numpy(clouds.clustered(3, n, n_clusters=10, domain=1.0e6, spread=100.0)).cuda()
    for impl in (0, 1):
        for minpts in (2, 5):
            timed("dbscan impl=%d minpts=%d" % (impl, minpts), lambda: abx.dbscan(space, xc, 200.0, minpts, abx.DBSCANParameters(impl, 0)), reps=1)


print(torch.cuda.get_device_name(0), "launches so far", abx.launch_count())
stage("sort", t_sort)
stage("bounds", t_bounds)
stage("codes", t_codes)
stage("tree structure", t_tree)
stage("spatial", t_spatial)
stage("nearest", t_nearest)
stage("half traversal", t_half)
stage("dbscan", t_dbscan)
if "--perf" in sys.argv:
    stage("perf 10M", t_perf)
print("launches", abx.launch_count())
