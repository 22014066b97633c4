#!/bin/bash
# MST / dendrogram / HDBSCAN parity on the GPU, the distributed tests with the shelved two-stage form, then an MST
# timing at 1M and 10M points
cd "$(dirname "$0")/.."
python -m pytest tests/test_mst.py tests/test_distributed.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import time, numpy as np, torch
import arborx_b200 as abx
from tests import clouds
space = abx.ExecutionSpace()
for n in (1_000_000, 10_000_000):
    for kind in ("uniform", "gantao"):
        pts = clouds.filled_box(0x5EED0001, n) if kind == "uniform" else clouds.gan_tao(3, n)
        d = torch.from_numpy(np.ascontiguousarray(pts, np.float32)).cuda()
        for k in (1, 5):
            abx.profile_enable(True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mst = abx.MinimumSpanningTree(space, d, k)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            prof = abx.profile_report()
            abx.profile_enable(False)
            print("MST n=%d %s k=%d: %.1f ms, %d rounds, total weight %.6g" % (n, kind, k, dt * 1e3, mst.iterations, float(mst.weights.double().sum())), flush=True)
            for name, cnt, ms, mx in prof[:4]:
                print("     %-40s %4d launches %9.3f ms (max %.3f)" % (name, cnt, ms, mx))
PY
