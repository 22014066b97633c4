#!/bin/bash
# final single-GPU validation + ncu evidence for the MST kernels + launch list of the bench command
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/r02_call27.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'componentNearestKernel|reduceLabelsKernel' \
  -s 2 -c 8 -f -o gpurun_out/prof_r02_mst python scripts/profile_mst.py 10000000 1 > gpurun_out/r02_ncu_mst.log 2>&1
echo "mst capture rc=$?"; tail -2 gpurun_out/r02_ncu_mst.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_final.csv \
  python bench.py --steps 2 --warmup 1 --skip-workloads --e2e-steps 1 > gpurun_out/r02_ncu_bench_final.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/r02_launches_final.csv | cut -c1-200
ls -la gpurun_out/prof_r02_mst.ncu-rep
