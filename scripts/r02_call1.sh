#!/bin/bash
# round 2, GPU call 1: wide-node validation + ncu captures of the DBSCAN main kernels at 10M
set -u
cd "$(dirname "$0")/.."
export ABX_LIBRARY=$PWD/arborx_b200/lib/libabx_tuning.so
mkdir -p gpurun_out
bash scripts/validate_wide.sh 2>&1 | tee gpurun_out/r02_validate_wide.log
for impl in 1 0; do
  timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'denseMainKernel|fdbscanMainKernel|denseCountKernel|countCoreKernel' -s 2 -c 2 \
    -f -o gpurun_out/prof_r02_dbscan_impl$impl python scripts/profile_dbscan.py 10000000 $impl > gpurun_out/ncu_dbscan_$impl.log 2>&1
  tail -2 gpurun_out/ncu_dbscan_$impl.log
done
