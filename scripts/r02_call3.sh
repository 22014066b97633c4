#!/bin/bash
# round 2, GPU call 3: the new bench line (headline + dbscan / triangles / 100M workloads) and the reference arm
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read())
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
c = d["components"]
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in c.items() if not k.endswith("note")})
for k, w in d["workloads"].items():
    print(k, json.dumps(w)[:1500])
print(d["cpu_baseline"])
print(json.loads(open("gpurun_out/r02_bench_ref.json").read())["value"])
PY
