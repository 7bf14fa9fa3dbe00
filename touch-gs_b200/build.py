"""Build libtgs.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python touch-gs_b200/build.py [--force]

The built ``libtgs.so`` is git-ignored but travels to the GPU box with the gpurun snapshot.
``preprocess.cu`` is compiled with ``--fmad=false`` (bit-exact integer stage).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libtgs.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
          "--expt-relaxed-constexpr", "-I", os.path.join(HERE, "..", "include")]
UNITS = {
    "preprocess.cu": ["--fmad=false"],
    "binning.cu": [],
    "render.cu": [],
    "refstructure.cu": [],
    "train_ops.cu": [],
    "screen_api.cu": [],
    "api.cu": [],
    "host_step.cu": [],
    "touch_inputs.cu": ["--fmad=false"],
}


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(srcs, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "tgs.h"))
    headers.append(os.path.abspath(__file__))
    units = {u: f for u, f in UNITS.items() if os.path.exists(os.path.join(CSRC, u))}

    def compile_one(item):
        unit, flags = item
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        if force or _stale([src] + headers, obj):
            cmd = [nvcc, *ARCH, *COMMON, *flags, "-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = os.path.join(OBJ, unit.replace(".cu", ".ptxas.log"))
            with open(log, "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {unit}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(units)) as ex:
        objs = list(ex.map(compile_one, units.items()))
    if force or _stale(objs, OUT):
        cmd = [nvcc, *ARCH, "-shared", "-o", OUT, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
