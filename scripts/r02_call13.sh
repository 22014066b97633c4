#!/bin/bash
# round 2, GPU call 13: ncu evidence -- launch list of the bench command, full captures of the step's kernels and of
# the reworked DenseBox kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --skip-workloads --e2e-steps 1 > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/r02_launches.csv | cut -c1-200
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'nearestKernel|spatialKernel|hierarchyLocalKernel|wideConvertKernel|onesweepPassKernel' -s 14 -c 14 \
  -f -o gpurun_out/prof_r02_step python scripts/profile_kernels.py > gpurun_out/r02_ncu_step.log 2>&1
echo "step capture rc=$?"; tail -2 gpurun_out/r02_ncu_step.log
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'denseCellPairsKernel|sparseMainKernel|denseCountKernel|finalizeLabelsKernel|denseCellUnionKernel' -s 5 -c 5 \
  -f -o gpurun_out/prof_r02_densebox python scripts/profile_dbscan.py 10000000 1 > gpurun_out/r02_ncu_densebox.log 2>&1
echo "densebox capture rc=$?"; tail -2 gpurun_out/r02_ncu_densebox.log
ls -la gpurun_out/*.ncu-rep | tail -4
