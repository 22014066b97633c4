#!/bin/bash
# round 2, GPU call 14: round-free chunk hierarchy kernel -- parity suite, build profile, headline bench; PCIe probe
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02_pytest_call14.log
timeout 300 python scripts/profile_build_big.py 10000000 2>&1 | tee gpurun_out/r02_build_chunk.log
timeout 600 python bench.py --steps 10 --warmup 3 --skip-workloads > gpurun_out/r02_bench_c14.json 2> gpurun_out/r02_bench_c14.err
python - <<'PY'
import json, torch, time
d = json.loads(open("gpurun_out/r02_bench_c14.json").read())
c = d["components"]
print("step", round(d["ms_per_step"], 3), "build/radius/knn", round(c["build_ms"], 3), round(c["radius_ms"], 3), round(c["knn_ms"], 3), "e2e", round(d["e2e"]["ms_per_step"], 2))
# pinned-memory copy rates of this box (what bounds the end-to-end path)
n = 1 << 28
h = torch.empty(n, dtype=torch.int32).pin_memory()
g = torch.empty(n, dtype=torch.int32, device="cuda")
for name, fn in (("H2D", lambda: g.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(g, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(name, "GB/s", 3 * n * 4 / (time.perf_counter() - t0) / 1e9)
s2 = torch.cuda.Stream()
t0 = time.perf_counter()
for _ in range(3):
    g.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2 = h  # same pinned buffer is fine for a rate probe
        g2 = torch.empty_like(g) if _ == 0 else g2
        h2.copy_(g2, non_blocking=True)
torch.cuda.synchronize()
print("H2D + D2H concurrently: GB/s per direction", 3 * n * 4 / (time.perf_counter() - t0) / 1e9)
PY
