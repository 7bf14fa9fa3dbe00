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


EXT_SRC = os.path.join(CSRC, "torch_ext.cpp")
EXT_OUT = os.path.join(HERE, "_C.so")


def build_torch_ext(force: bool = False) -> str:
    """g++ build of the PyTorch binding `_C` (csrc/torch_ext.cpp) against the installed torch and the in-tree
    libtgs.so (found at run time through an $ORIGIN rpath).  No CUDA code lives here: nothing to compile with nvcc."""
    import torch
    from torch.utils import cpp_extension as ce
    lib = build()
    deps = [EXT_SRC, os.path.join(HERE, "..", "include", "tgs.h"), os.path.abspath(__file__), lib]
    if not force and not _stale(deps, EXT_OUT):
        return EXT_OUT
    inc = []
    for d in ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True):
        inc += ["-isystem", d]
    import sysconfig
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(getattr(torch._C, "_GLIBCXX_USE_CXX11_ABI", 1))
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=touchgs_b200_C",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc, EXT_SRC, "-o", EXT_OUT,
           f"-L{HERE}", "-l:libtgs.so", f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
           "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(OBJ, "torch_ext.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for torch_ext.cpp:\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}")
    return EXT_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_ext(force="--force" in sys.argv))
