"""The C++ facade (include/ArborX_B200.hpp) compiles with a plain host compiler against the
C ABI (CPU suite) and reproduces the reference's known answers on the GPU (-m gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "arborx_b200", "lib", "facade_example")


def _compile():
    lib = os.path.join(ROOT, "arborx_b200", "lib")
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "facade_example.cpp"), "-o", EXE, "-L" + lib, "-labx",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + lib, "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)


def test_facade_compiles():
    from arborx_b200 import _lib
    _lib.lib()
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_facade_runs():
    if not os.path.exists(EXE):
        _compile()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "FACADE OK" in out.stdout
    assert "offsets: 0 1 4 8" in out.stdout
