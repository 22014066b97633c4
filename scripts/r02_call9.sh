#!/bin/bash
# round 2, GPU call 9 (1 GPU): the distributed tests (in-process ranks) with the side-stream exchange, full suite
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02_pytest_call9.log
