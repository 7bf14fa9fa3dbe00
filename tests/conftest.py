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


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without CUDA: gpu-marked tests are skipped instead of failing (the CPU suite is
    then green without `-m "not gpu"`).  On a GPU box nothing is skipped: the product path has no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Float-parity report: for every assert_close_tensor call, the measured rel_inf and which branch passed."""
    try:
        import helpers
    except Exception:  # noqa: BLE001
        return
    log = helpers.PARITY_LOG
    if not log:
        return
    import json
    used = [e for e in log if e["branch"] != "rel_inf"]
    tr = terminalreporter
    tr.write_sep("-", f"float parity: {len(log)} tensor comparisons, {len(log) - len(used)} within rel_inf, "
                      f"{len(used)} needed the outlier budget")
    worst = max(log, key=lambda e: e["rel_inf"] if e["branch"] == "rel_inf" else 0.0)
    tr.write_line(f"largest rel_inf that passed on its own: {worst['rel_inf']:.3e} ({worst['test']} :: {worst['tensor']})")
    for e in used:
        tr.write_line(f"  {e['branch']:>14}  rel_inf={e['rel_inf']:.3e}  outliers={e['outlier_frac']:.3e} "
                      f"(budget {e.get('budget', 0):g})  {e['test']} :: {e['tensor']}")
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(log, f, indent=1)
    except OSError:
        pass


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
