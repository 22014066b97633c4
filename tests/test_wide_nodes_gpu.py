"""The experimental 4-wide quantised nodes (ABX_WIDE): the spatial and kNN parity tests re-run in a
subprocess with the wide walk switched on (the switch is read once per process)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level,select", [
    ("1", "spatial_sphere_vs_oracle or spatial_box_and_point or spatial_vs_bruteforce or spatial_boundary"),
    ("2", "nearest_vs_oracle or nearest_bruteforce or spatial_sphere_vs_oracle"),
])
def test_parity_with_wide_nodes(level, select):
    env = dict(os.environ, ABX_WIDE=level)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_gpu.py"), "-m", "gpu",
                          "-q", "-x", "-k", select, "-p", "no:cacheprovider"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout
