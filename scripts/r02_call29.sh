#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests/test_triangles_gpu.py tests/test_rays.py tests/test_facade.py tests/test_callbacks_oracle.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -12
python scripts/debug_rays.py 2>&1 | head -4
