#!/bin/bash
# round 2 validation on one B200: smoke, the whole GPU suite, the bench line (default flags), the reference arm
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_final.log
timeout 1500 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1_final.err
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference_final.json 2> gpurun_out/r02_bench_reference_final.err
echo "reference rc=$?"; cut -c1-700 gpurun_out/r02_bench_reference_final.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_final.json").read())
c = d["components"]
print("N=1:", round(d["ms_per_step"], 3), round(d["value"], 1), "build/radius/knn", round(c["build_ms"], 3), round(c["radius_ms"], 3), round(c["knn_ms"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), round(d["e2e"]["value"], 1))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, w in d["workloads"].items():
    print(k, json.dumps(w)[:1500])
print(d["roofline"])
PY
