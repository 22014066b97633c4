"""Device callbacks (include/ArborX_B200_Callbacks.cuh) against the CPU oracle -- not against the library's own CRS
results: attach(predicates, data), the PerThread single query issued from a user kernel, ordered_intersects(ray)
with and without early exit (spatial/detail/ArborX_TreeTraversal.hpp:338-489)."""
import os
import subprocess

import numpy as np
import pytest

from tests import clouds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "arborx_b200", "lib", "callback_check")
F = np.float32
CAP = 64


def _compile():
    lib = os.path.join(ROOT, "arborx_b200", "lib")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17",
                           "--extended-lambda", "--expt-relaxed-constexpr", "-fmad=false", "-O2",
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "callback_check.cu"),
                           "-o", EXE, "-L" + lib, "-labx", "-Xlinker", "-rpath", "-Xlinker", lib])


def test_callback_check_compiles():
    from arborx_b200 import _lib
    _lib.lib()
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [(1, 1), (1, 2), (1, 3000), (0, 3000), (1, 60_000)])
def test_callbacks_against_oracle(tmp_path, kind, n):
    import oracle
    if not os.path.exists(EXE):
        _compile()
    lo = clouds.filled_box(91, max(n, 8))[:n]
    prims = lo if kind == 0 else np.concatenate([lo, lo + clouds.uniform01(92, max(n, 8))[:n] * F(1.2)], 1).astype(F)
    qs, qr = 1500, 1500
    a = F(np.cbrt(float(max(n, 8))))
    c = clouds.filled_box(93, qs) * (a / F(np.cbrt(float(qs))))
    spheres = np.concatenate([c, np.full((qs, 1), 1.7, F)], 1).astype(F)
    rays = (clouds.ball_rays(94, qr) * np.array([a, a, a, 1, 1, 1], F)).astype(F)
    rays[::17, 3:] = np.array([0, 1, 0], F)
    tags = np.random.default_rng(5).permutation(qs).astype(np.int32)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([n, kind, qs, qr], np.int32).tofile(f)
        prims.astype(F).tofile(f)
        spheres.tofile(f)
        rays.tofile(f)
        tags.tofile(f)
    out = subprocess.run([EXE, fin, fout], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "CALLBACK CHECK WRITTEN" in out.stdout, out.stdout + out.stderr
    raw = np.fromfile(fout, np.int32)
    p = 0
    attach = raw[p:p + qs]; p += qs
    per_thread = raw[p:p + qs]; p += qs
    ocount = raw[p:p + qr]; p += qr
    ovals = raw[p:p + qr * CAP].view(np.uint32).reshape(qr, CAP); p += qr * CAP
    odist = raw[p:p + qr * CAP].view(F).reshape(qr, CAP); p += qr * CAP
    fval = raw[p:p + qr]; p += qr
    fdist = raw[p:p + qr].view(F); p += qr
    assert p == raw.size

    tree = oracle.Tree(prims, kind)
    counts = tree.spatial_count(spheres)
    # attach: the callback saw data[query] = its tag
    expect = np.zeros(qs, np.int32)
    expect[tags] = counts
    assert np.array_equal(attach, expect)
    assert np.array_equal(per_thread, counts)
    # ordered rays: the same leaves with the same entry distances, handed out nearest first
    roff, ridx, rd = tree.ordered_ray_crs(rays)
    assert np.array_equal(ocount, np.diff(roff))
    assert int(ocount.max()) <= CAP
    inf = np.float32(np.inf)
    for i in range(qr):
        m = ocount[i]
        gv, gd = ovals[i, :m], odist[i, :m]
        ev, ed = ridx[roff[i]:roff[i + 1]], rd[roff[i]:roff[i + 1]]
        finite = gd[gd < inf]
        assert np.all(np.diff(finite) >= 0)  # nearest first
        assert np.array_equal(np.sort(gd), np.sort(ed))
        assert sorted(zip(gd.tolist(), gv.tolist())) == sorted(zip(ed.tolist(), ev.tolist()))
    # early exit: the first callback of every ray
    foff, fidx, fd = tree.ordered_ray_crs(rays, limit=1)
    has = np.diff(foff) > 0
    assert np.array_equal(fval >= 0, has)
    assert np.array_equal(fdist[has], fd)
    for i in np.nonzero(has)[0]:
        first_d = rd[roff[i]]
        cands = ridx[roff[i]:roff[i + 1]][rd[roff[i]:roff[i + 1]] == first_d]
        assert fval[i] in cands
