#!/bin/bash
# round 2, GPU call 6: chunked kNN kernel (tuning library): parity at chunk 64, timings at 0 / 64 / 96 / 128
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export ABX_LIBRARY=$PWD/arborx_b200/lib/libabx_tuning.so
echo "== kNN parity, ABX_KNN_CHUNK=64"
ABX_KNN_CHUNK=64 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_dist_kernels_gpu.py tests/test_distributed.py tests/test_full_size_gpu.py -m gpu -q -x -k "nearest or knn or distributed or golden" 2>&1 | tail -5
for ch in 0 64 96 128; do
  ABX_KNN_CHUNK=$ch timeout 600 python bench.py --steps 10 --warmup 3 --skip-workloads --e2e-steps 1 > gpurun_out/r02_bench_c6_$ch.json 2> gpurun_out/r02_bench_c6.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_c6_$ch.json").read())
c = d["components"]
print("chunk $ch:", round(d["ms_per_step"], 3), round(c["build_ms"], 3), round(c["radius_ms"], 3), round(c["knn_ms"], 3))
for k in d["kernels"][:3]:
    print("  ", k["kernel"], k["launches"], k["avg_ms"])
PY
done
