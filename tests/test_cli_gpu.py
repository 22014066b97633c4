"""scripts/bvh_driver.py: the reference's bvh_driver command line and Google-Benchmark-shaped output
(benchmarks/bvh_driver/benchmark_registration.hpp:91-126) as scripts/benchmark.py parses it."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests import clouds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_point_cloud_kinds():
    n = 6000
    a = np.cbrt(float(n))
    fb = clouds.point_cloud("filled_box", 3, n)
    hb = clouds.point_cloud("hollow_box", 3, n)
    fs = clouds.point_cloud("filled_sphere", 3, n)
    hs = clouds.point_cloud("hollow_sphere", 3, n)
    assert np.abs(fb).max() <= a
    # hollow box: point i sits on the face of axis (i / 2) % 3, side i % 2 (PointClouds.hpp:101-123)
    i = np.arange(n)
    assert np.allclose(hb[i, (i // 2) % 3], np.where(i % 2 == 0, -a, a))
    assert np.linalg.norm(fs, axis=1).max() <= a * 1.0001
    assert np.allclose(np.linalg.norm(hs, axis=1), a, rtol=1e-5)


@pytest.mark.gpu
def test_bvh_driver_cli():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bvh_driver.py"), "--exact-spec",
                          "20000/10000/10/1/0/0/2", "--exact-spec", "5000/5000/5/1/0/1/3", "--repetitions", "3"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    data = json.loads(out.stdout[out.stdout.index("{"):])["benchmarks"]
    names = [b["name"] for b in data]
    # the templates of scripts/benchmark.py:55-66
    assert any(re.search(r"BM_construction<.*B200[^/]*/20000/0/manual_time_median", x) for x in names)
    assert any(re.search(r"BM_knn_search<.*B200[^/]*/20000/10000/[^/]*/1/0/", x) for x in names)
    assert any(re.search(r"BM_radius_search<.*B200[^/]*/5000/5000/[^/]*/1/0/1/", x) for x in names)
    for b in data:
        if b["aggregate_name"] == "median":
            assert b["rate"] > 0


@pytest.mark.gpu
def test_dbscan_driver_on_the_benchmark_input(tmp_path):
    """ArborX_Benchmark_DBSCAN --filename=input.txt --eps=1.4 --verify (benchmarks/cluster/CMakeLists.txt:13) on the
    benchmark's own 8-point file."""
    import subprocess
    import sys
    txt = tmp_path / "input.txt"
    txt.write_text("8 3\n0 0 0\n1 1 1\n2 2 2\n3 3 3\n9 9 9\n10 10 10\n11 11 11\n20 20 20\n")
    for impl in ("fdbscan", "fdbscan-densebox"):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "dbscan_driver.py"), "--filename", str(txt),
                              "--eps", "1.8", "--impl", impl, "--verify"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "#clusters       : 2" in out.stdout and "Verification passed" in out.stdout
        assert "#noise   points : 1 " in out.stdout
