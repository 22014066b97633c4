#!/bin/bash
# round 2, GPU call 11: new features (nearest geometries, indexed triangles, callbacks vs oracle) + full suite + build profile at 100M
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r02_pytest_call11.log
timeout 600 python scripts/profile_build_big.py 10000000 50000000 100000000 2>&1 | tee gpurun_out/r02_build_big.log
