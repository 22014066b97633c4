#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_mst.py tests/test_facade.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import time, numpy as np, torch
import arborx_b200 as abx
from tests import clouds
space = abx.ExecutionSpace()
n = 10_000_000
d = torch.from_numpy(np.ascontiguousarray(clouds.gan_tao(3, n), np.float32)).cuda()
for impl, name in ((abx.DENDROGRAM_BORUVKA, "boruvka"), (abx.DENDROGRAM_UNION_FIND, "union_find")):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = abx.hdbscan(space, d, 5, impl)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("hdbscan 10M gantao core_min_size=5 %s: %.1f ms" % (name, dt * 1e3), flush=True)
PY
