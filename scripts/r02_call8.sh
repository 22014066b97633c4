#!/bin/bash
# round 2, GPU call 8 (N GPUs, default 2): C++ DistributedTree over NCCL -- correctness cases, then the bench line
set -u
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR scripts/dist_check.py 2>&1 | tail -15 | tee gpurun_out/r02_dist_check_n$N.log
timeout 1500 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n$N.json").read())
c = d["components"]
print("N=$N:", round(d["ms_per_step"], 3), round(d["value"], 1), "build/radius/knn", round(c["build_ms"], 3), round(c["radius_ms"], 3), round(c["knn_ms"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2), round(d["e2e"]["value"], 1))
print("d2h", d["e2e"]["d2h_bytes_per_step"], d["e2e"]["api"])
for k in d["kernels"][:14]:
    print("  ", k["kernel"], k["launches"], k["avg_ms"], k["max_ms"])
for k, w in d["workloads"].items():
    print(k, json.dumps(w)[:900])
print(d["roofline"])
PY
