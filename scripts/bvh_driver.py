"""bvh_driver-shaped command line for the B200 library (benchmarks/bvh_driver/bvh_driver.cpp,
benchmark_registration.hpp:56-126, 219-400): BM_construction / BM_radius_search / BM_knn_search over the four
point-cloud kinds, Google-Benchmark-shaped JSON on stdout (names and the `rate` counter as the reference's
scripts/benchmark.py parses them: "BM_construction<ArborX::BVH<B200>>/n/cloud/manual_time_median", ...).

    python scripts/bvh_driver.py --values 1000000 --queries 1000000 --neighbors 10 --buffer 0 \
        --source-point-cloud-type filled_box --target-point-cloud-type filled_sphere --repetitions 5
    python scripts/bvh_driver.py --exact-spec 10000000/10000000/10/1/0/0/2    # n/q/k/sort/buffer/source/target

Times are CUDA-event times on the launching stream (Google Benchmark's manual time); rate = items per second
(primitives for construction, queries for the searches)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from tests import clouds  # noqa: E402

KINDS = ["filled_box", "hollow_box", "filled_sphere", "hollow_sphere"]
TREE = "ArborX::BVH<B200>"


def timed(fn, reps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return np.array(ts)


def entries(name, ts, items):
    out = []
    for agg, t in (("mean", ts.mean()), ("median", float(np.median(ts))), ("stddev", ts.std())):
        e = {"name": "%s/manual_time_%s" % (name, agg), "run_name": name + "/manual_time", "run_type": "aggregate",
             "aggregate_name": agg, "repetitions": len(ts), "real_time": t * 1e9, "cpu_time": t * 1e9,
             "time_unit": "ns"}
        if agg != "stddev":
            e["rate"] = items / t
        out.append(e)
    return out


def run_spec(n, q, k, sort, buffer, source, target, reps):
    space = abx.ExecutionSpace()
    pts = torch.from_numpy(clouds.point_cloud(KINDS[source], 0x5EED0001, n)).cuda()
    # queries: the target cloud at the same scale (makeSpatialQueries / makeNearestQueries, :163-215)
    qp = clouds.point_cloud(KINDS[target], 0x5EED0002, q) * np.float32(np.cbrt(float(n)) / np.cbrt(float(q)))
    qpts = torch.from_numpy(qp.astype(np.float32)).cuda()
    r = float(clouds.bvh_driver_radius(k))
    spheres = torch.cat([qpts, torch.full((q, 1), r, device="cuda")], 1).contiguous()
    policy = abx.TraversalPolicy().setPredicateSorting(bool(sort)).setBufferSize(buffer)
    out = []
    holder = {}

    def build():
        holder["bvh"] = abx.BoundingVolumeHierarchy(space, pts)

    out += entries("BM_construction<%s>/%d/%d" % (TREE, n, source), timed(build, reps), n)
    bvh = holder["bvh"]
    p_sp, p_nn = abx.intersects(spheres), abx.nearest(qpts, k)
    out += entries("BM_radius_search<%s>/%d/%d/%d/%d/%d/%d/%d" % (TREE, n, q, k, sort, buffer, source, target),
                   timed(lambda: bvh.query(space, p_sp, policy), reps), q)
    out += entries("BM_knn_search<%s>/%d/%d/%d/%d/%d/%d" % (TREE, n, q, k, sort, source, target),
                   timed(lambda: bvh.query(space, p_nn, policy), reps), q)
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--values", type=int, default=50000)
    ap.add_argument("--queries", type=int, default=20000)
    ap.add_argument("--neighbors", type=int, default=10)
    ap.add_argument("--buffer", type=int, default=0)
    ap.add_argument("--predicate-sort", type=int, default=1)
    ap.add_argument("--source-point-cloud-type", default="filled_box", choices=KINDS)
    ap.add_argument("--target-point-cloud-type", default="filled_sphere", choices=KINDS)
    ap.add_argument("--exact-spec", action="append", default=[],
                    help="n_values/n_queries/n_neighbors/sort_predicates/buffer_size/source/target (repeatable)")
    ap.add_argument("--repetitions", type=int, default=5)
    a = ap.parse_args()
    specs = [tuple(int(x) for x in s.split("/")) for s in a.exact_spec] or [
        (a.values, a.queries, a.neighbors, a.predicate_sort, a.buffer, KINDS.index(a.source_point_cloud_type),
         KINDS.index(a.target_point_cloud_type))]
    benchmarks = []
    for spec in specs:
        benchmarks += run_spec(*spec, reps=a.repetitions)
    props = torch.cuda.get_device_properties(0)
    print(json.dumps({"context": {"executable": "scripts/bvh_driver.py", "library": "libabx.so (sm_100a)",
                                  "device": props.name, "num_sms": props.multi_processor_count},
                      "benchmarks": benchmarks}, indent=1))


if __name__ == "__main__":
    main()
