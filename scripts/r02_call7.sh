#!/bin/bash
# round 2, GPU call 7: lazy wide conversion + kNN candidate-distance / stack placement variants (tuning library)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02_pytest_call7.log
export ABX_LIBRARY=$PWD/arborx_b200/lib/libabx_tuning.so
for v in 1 2; do
  echo "== kNN parity, ABX_KNN_DREG=$v"
  ABX_KNN_DREG=$v timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_dist_kernels_gpu.py -m gpu -q -x -k "nearest or knn or golden" 2>&1 | tail -3
done
for v in 0 1 2; do
  ABX_KNN_DREG=$v timeout 600 python bench.py --steps 10 --warmup 3 --skip-workloads --e2e-steps 1 > gpurun_out/r02_bench_c7_$v.json 2> gpurun_out/r02_bench_c7.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_c7_$v.json").read())
c = d["components"]
print("dreg $v:", round(d["ms_per_step"], 3), round(c["build_ms"], 3), round(c["radius_ms"], 3), round(c["knn_ms"], 3))
for k in d["kernels"][:3]:
    print("  ", k["kernel"], k["launches"], k["avg_ms"])
PY
done
