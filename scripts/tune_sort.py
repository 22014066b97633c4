"""Sort tuning aid: correctness + timing of abx_sort_u64/u32 for the tile shape picked by ABX_SORT_CONFIG."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import arborx_b200 as abx
from arborx_b200 import _lib
L = _lib.lib()
space = abx.ExecutionSpace()
cfg = os.environ.get("ABX_SORT_CONFIG", "default")
ok = True
for n in (1000, 4097, 1_000_003):
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    k = torch.randint(0, 2**62, (n,), device="cuda", dtype=torch.int64, generator=g)
    k[::5] = k[0]
    ref_k, ref_p = torch.sort(k, stable=True)
    kk = k.clone(); perm = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.check(L.abx_sort_u64(space.handle, C.c_void_p(kk.data_ptr()), C.c_void_p(perm.data_ptr()), n))
    ok &= bool(torch.equal(kk, ref_k)) and bool(torch.equal(perm.long(), ref_p))
    k32 = torch.randint(0, 2**30, (n,), device="cuda", dtype=torch.int32, generator=g)
    ref_k, ref_p = torch.sort(k32, stable=True)
    kk = k32.clone()
    _lib.check(L.abx_sort_u32(space.handle, C.c_void_p(kk.data_ptr()), C.c_void_p(perm.data_ptr()), n))
    ok &= bool(torch.equal(kk, ref_k)) and bool(torch.equal(perm.long(), ref_p))
n = 10_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
k = torch.randint(0, 2**62, (n,), device="cuda", dtype=torch.int64, generator=g)
k32 = torch.randint(0, 2**30, (n,), device="cuda", dtype=torch.int32, generator=g)
perm = torch.empty(n, dtype=torch.int32, device="cuda")
def timeit(fn, src):
    best = 1e9
    for _ in range(6):
        kk = src.clone(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(kk); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t64 = timeit(lambda kk: _lib.check(L.abx_sort_u64(space.handle, C.c_void_p(kk.data_ptr()), C.c_void_p(perm.data_ptr()), n)), k)
t32 = timeit(lambda kk: _lib.check(L.abx_sort_u32(space.handle, C.c_void_p(kk.data_ptr()), C.c_void_p(perm.data_ptr()), n)), k32)
print("config %s: correct=%s  sort_u64 10M: %.3f ms  sort_u32 10M: %.3f ms" % (cfg, ok, t64, t32), flush=True)
