#!/bin/bash
# round 2, GPU call 4: GPU suite with the DenseBox rework + new full-size oracle tests, then the bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/r02_pytest_call4.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1_b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_b.json").read())
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
w = d["workloads"]["dbscan_10M"]
for k, v in w.items():
    if isinstance(v, dict):
        print(k, v.get("ms"), v.get("kernels_ms_per_call"))
PY
