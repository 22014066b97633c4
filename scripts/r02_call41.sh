#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_run.py 20000 > gpurun_out/r02_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitize_memcheck.log
timeout 1500 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n1_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_final.json").read())
c = d["components"]
print("N=1:", round(d["ms_per_step"], 3), round(d["value"], 1), "build/radius/knn", round(c["build_ms"], 3), round(c["radius_ms"], 3), round(c["knn_ms"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), round(d["e2e"]["value"], 1))
w = d["workloads"]
print(json.dumps(w["mst_10M"])[:1800])
t = w["triangles_20M"]
print({k: (round(v["ms"], 2), [b for a, b in v.items() if "match" in a]) for k, v in t.items() if isinstance(v, dict)})
print({k: round(v["ms"], 2) for k, v in w["dbscan_10M"].items() if isinstance(v, dict)})
PY
