#!/bin/bash
# round 2, GPU call 16: C++ distributed DBSCAN (in-process ranks) + full suite
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r02_pytest_call16.log
