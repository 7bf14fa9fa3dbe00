import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def host_math_lib():
    """g++ build of the __host__ __device__ per-Gaussian math (tests/host_math_harness.cpp)."""
    import ctypes
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostmath.so")
    src = os.path.join(ROOT, "tests", "host_math_harness.cpp")
    hdrs = [os.path.join(ROOT, "touch-gs_b200", "csrc", h) for h in ("tgs_math.cuh", "touch_inputs_math.cuh")]
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def tgs_lib():
    """libtgs.so, built on demand (nvcc cross-compiles without a GPU)."""
    import importlib
    build = importlib.import_module("touch-gs_b200.build")
    build.build()
    import touchgs_b200
    return touchgs_b200._lib.load()
