#!/bin/bash
# round 2, GPU call 15: chunk hierarchy kernel variants (tuning library): parity + build profile for each
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export ABX_LIBRARY=$PWD/arborx_b200/lib/libabx_tuning.so
for v in 0 1 2 3 4; do
  echo "== ABX_HIER_CHUNK=$v"
  ABX_HIER_CHUNK=$v timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q -x -k "tree or build or structure or golden or degenerate or chain or duplicat" 2>&1 | tail -2
  ABX_HIER_CHUNK=$v timeout 300 python scripts/profile_build_big.py 10000000 2>&1 | head -5
done
