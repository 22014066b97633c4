"""ArborX_Benchmark_DBSCAN-shaped command line (benchmarks/cluster/dbscan.cpp:196-232, CMakeLists.txt:13):
    python scripts/dbscan_driver.py --filename input.txt --eps 1.4 --verify            # text file
    python scripts/dbscan_driver.py --filename pts.arborx --binary --eps 200 --core-min-size 5 --impl fdbscan-densebox
    python scripts/dbscan_driver.py --n 10000000 --eps 200 --core-min-size 5            # GanTao generator
Prints the reference's report (phase times from the library's per-kernel CUDA events, cluster / noise counts);
--verify runs the reference's verifier (restated in the oracle: test infrastructure, CPU, O(n * neighbours))."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx  # noqa: E402
from arborx_b200 import io  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--filename", default="")
ap.add_argument("--binary", action="store_true")
ap.add_argument("--max-num-points", type=int, default=-1)
ap.add_argument("--n", type=int, default=1_000_000, help="GanTao generator size when no file is given")
ap.add_argument("--eps", type=float, required=True)
ap.add_argument("--core-min-size", type=int, default=2)
ap.add_argument("--cluster-min-size", type=int, default=1)
ap.add_argument("--impl", default="fdbscan-densebox", choices=["fdbscan", "fdbscan-densebox"])
ap.add_argument("--algorithm", default="dbscan", choices=["dbscan", "dbscan*"])
ap.add_argument("--verify", action="store_true")
args = ap.parse_args()

if args.filename:
    pts = io.load_points(args.filename, args.binary, args.max_num_points)
    print('read in %d 3D points from "%s"' % (len(pts), args.filename))
else:
    from tests import clouds
    pts = clouds.gan_tao(3, args.n)
print("eps               : %f" % args.eps)
print("minpts            : %d" % args.core_min_size)
print("cluster min size  : %d" % args.cluster_min_size)
print("implementation    : %s" % args.impl)
print("algorithm         : %s" % args.algorithm)
space = abx.ExecutionSpace()
d = torch.from_numpy(pts).cuda()
params = abx.DBSCANParameters(1 if args.impl == "fdbscan-densebox" else 0, 1 if args.algorithm == "dbscan*" else 0)
abx.dbscan(space, d, args.eps, args.core_min_size, params)  # warm-up
torch.cuda.synchronize()
abx.profile_enable(True)
t0 = time.perf_counter()
labels = abx.dbscan(space, d, args.eps, args.core_min_size, params)
torch.cuda.synchronize()
total = time.perf_counter() - t0
prof = abx.profile_report()
abx.profile_enable(False)
ms = lambda names: sum(t for k, c, t, mx in prof if any(nm in k for nm in names)) * 1e-3
tp = time.perf_counter()
lab = labels.cpu().numpy()
ids, counts = np.unique(lab[lab >= 0], return_counts=True)
keep = ids[counts >= args.cluster_min_size]
in_cluster = np.isin(lab, keep)
post = time.perf_counter() - tp
if args.impl == "fdbscan-densebox":
    print("-- dense cells      : %10.3f" % ms(["cellIndices", "cellStart", "cellOffsets", "cellId", "cellClass", "reorderCells",
                                               "denseCellUnion", "mixedPrimitives", "gatherDensePoints"]))
print("-- construction     : %10.3f" % ms(["sceneBounds", "morton64", "onesweep", "radix", "segmentFix", "hierarchy",
                                           "prefixSample"]))
print("-- query+cluster    : %10.3f" % ms(["countCore", "fdbscanMain", "denseCount", "denseCellPairs", "sparseMain",
                                           "finalizeLabels", "markNoise", "markDenseCore"]))
print("-- postprocess      : %10.3f" % post)
print("total time          : %10.3f" % total)
n = len(pts)
print("\n#clusters       : %d" % len(keep))
print("#cluster points : %d [%.2f%%]" % (int(in_cluster.sum()), 100.0 * in_cluster.sum() / max(n, 1)))
print("#noise   points : %d [%.2f%%]" % (n - int(in_cluster.sum()), 100.0 * (n - in_cluster.sum()) / max(n, 1)))
if args.verify:
    import oracle
    ok = oracle.dbscan_verify(pts, args.eps, args.core_min_size, lab, 1 if args.algorithm == "dbscan*" else 0) == 0
    print("Verification %s" % ("passed" if ok else "failed"))
    sys.exit(0 if ok else 1)
