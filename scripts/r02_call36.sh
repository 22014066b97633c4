#!/bin/bash
# e2e with / without NUMA pinning on the same box; distributed DBSCAN tests after the status agreement change
cd "$(dirname "$0")/.."
python -m pytest tests/test_distributed.py -m gpu -x -q -k "dbscan" 2>&1 | tail -3
for pin in 0 1 0 1; do
  if [ $pin = 0 ]; then export ABX_BENCH_NO_PIN=1; else unset ABX_BENCH_NO_PIN; fi
  python bench.py --skip-workloads --steps 5 --warmup 3 --cpu-n 100000 > gpurun_out/r02_bench_pin$pin.json 2> gpurun_out/r02_bench_pin$pin.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_pin$pin.json").read())
print("pin=$pin step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["e2e"]["api"][-20:], "cpu cores", d["cpu_baseline"]["cores"])
PY
done
nvidia-smi topo -m 2>/dev/null | head -12
