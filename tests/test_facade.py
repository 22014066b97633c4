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


CB_EXE = os.path.join(ROOT, "arborx_b200", "lib", "callback_example")


def _compile_callbacks():
    lib = os.path.join(ROOT, "arborx_b200", "lib")
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--extended-lambda",
           "--expt-relaxed-constexpr", "-fmad=false", "-O2", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "callback_example.cu"), "-o", CB_EXE, "-L" + lib, "-labx",
           "-Xlinker", "-rpath", "-Xlinker", lib]
    subprocess.check_call(cmd)


def test_callbacks_header_compiles():
    """include/ArborX_B200_Callbacks.cuh instantiates the traversal cores in a user translation unit."""
    from arborx_b200 import _lib
    _lib.lib()
    _compile_callbacks()
    assert os.path.exists(CB_EXE)


@pytest.mark.gpu
def test_callbacks_run():
    if not os.path.exists(CB_EXE):
        _compile_callbacks()
    out = subprocess.run([CB_EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "CALLBACKS OK" in out.stdout
