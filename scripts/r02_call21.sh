#!/bin/bash
# two-stage distributed kNN: parity with local ranks, then the whole distributed test file
python -m pytest tests/test_distributed.py tests/test_dist_kernels_gpu.py -m gpu -x -q 2>&1 | tail -15
