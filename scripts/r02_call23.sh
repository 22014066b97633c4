#!/bin/bash
# N GPUs: distributed check + bench, then the overlapped host timeline of the distributed queries (tuning library)
N=${1:-2}
cd "$(dirname "$0")/.."
bash scripts/r02_call8.sh $N
ABX_LIBRARY=$PWD/arborx_b200/lib/libabx_tuning.so ABX_DIST_TRACE=2 TRACE_STEPS=3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_trace.py > gpurun_out/r02_dist_trace2_n$N.log 2>&1
grep "abx trace" gpurun_out/r02_dist_trace2_n$N.log | tail -4
