#!/bin/bash
# round 2, GPU call 5: wide records by default (eager conversion, device-side fallback) -- suite + headline bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_call5.log
for cfg in "4 3" "8 4" "8 3"; do
  set -- $cfg
  timeout 600 python bench.py --steps 10 --warmup 3 --skip-workloads --e2e-chunks $1 --e2e-threads $2 > gpurun_out/r02_bench_c5_$1_$2.json 2> gpurun_out/r02_bench_c5.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_c5_$1_$2.json").read())
c = d["components"]
print("chunks $1 threads $2:", round(d["ms_per_step"], 3), round(c["build_ms"], 3), round(c["radius_ms"], 3), round(c["knn_ms"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2))
for k in d["kernels"][:8]:
    print("  ", k["kernel"], k["launches"], k["avg_ms"])
PY
done
