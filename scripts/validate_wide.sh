#!/bin/bash
# Round-2 checklist for the experimental 4-wide quantised nodes ("where the time is"):
#   gpurun --timeout 1500 -- 'bash scripts/validate_wide.sh'
# 1. the whole GPU suite with the wide walk in the spatial and the kNN kernels
# 2. the register-only conversion kernel against the same parity tests
# 3. bench lines: Node64 walk, wide spatial only, wide spatial + kNN, the same with the second conversion kernel
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== full GPU suite, ABX_WIDE=2"
ABX_WIDE=2 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== parity subset, ABX_WIDE=2 ABX_WIDE_CONVERT=2"
ABX_WIDE=2 ABX_WIDE_CONVERT=2 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q -x \
  -k "spatial or nearest or structured or buffer or unsorted or chain or duplicated or dbscan" 2>&1 | tail -3
for cfg in "0 1" "1 1" "2 1" "2 2"; do
  set -- $cfg
  echo "== bench ABX_WIDE=$1 ABX_WIDE_CONVERT=$2"
  ABX_WIDE=$1 ABX_WIDE_CONVERT=$2 ABX_BENCH_DEBUG=1 timeout 300 python bench.py --steps 6 --warmup 3 --cpu-sample 20000 \
    2>/dev/null | tail -1 > gpurun_out/bench_wide_$1_$2.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_wide_$1_$2.json").read())
print(d["ms_per_step"], d["components"]["build_ms"], d["components"]["radius_ms"], d["components"]["knn_ms"])
for k in d["kernels"][:6]:
    print("  ", k["kernel"], k["launches"], k["avg_ms"])
PY
done
