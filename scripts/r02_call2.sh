#!/bin/bash
# round 2, GPU call 2: full GPU suite on the release library (C++ DistributedTree, dist kernel tests, hygiene)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee gpurun_out/r02_pytest_call2.log
