#!/bin/bash
# N GPUs: the headline step with the distributed kNN split on / off (tuning library), no secondary workloads
N=${1:-8}
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for mode in 1 0; do
  ABX_LIBRARY=$PWD/arborx_b200/lib/libabx_tuning.so ABX_KNN_TWO_STAGE=$mode timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --skip-workloads --e2e-steps 1 > gpurun_out/r02_bench_split${mode}_n$N.json 2> gpurun_out/r02_bench_split${mode}_n$N.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_split${mode}_n$N.json").read())
c = d["components"]
print("split=$mode N=$N:", round(d["ms_per_step"], 3), round(d["value"], 1), "build/radius/knn", round(c["build_ms"], 3), round(c["radius_ms"], 3), round(c["knn_ms"], 3))
for k in d["kernels"][:4]:
    print("  ", k["kernel"], k["launches"], k["avg_ms"], k["max_ms"])
PY
done
