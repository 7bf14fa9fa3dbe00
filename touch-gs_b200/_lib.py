"""ctypes binding of libtgs.so (the C ABI declared in ``include/tgs.h``).

The library is the product: if it is missing or fails to load, importing the operator raises --
there is NO CPU fallback and nothing here ever imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TGS_LIB_PATH") or os.path.join(HERE, "libtgs.so")   # override: A/B builds

TGS_ABI_VERSION = 8
BUF_GEOM, BUF_BINNING, BUF_IMAGE, BUF_TEMP = 0, 1, 2, 3
LOSS_NONE, LOSS_L1, LOSS_L2 = 0, 1, 2
LOSS_MODES = {"none": LOSS_NONE, "l1": LOSS_L1, "l2": LOSS_L2}
NGRAD = 10


def screen_grad_floats(N: int, contrib_flags: bool = True) -> int:
    """float32 elements of a screen-gradient buffer (include/tgs.h: tgs_screen_grad_bytes): the [N,10] rows, plus --
    with TgsSettings.contrib_flags -- one contributor byte per Gaussian at the 128-byte aligned offset behind them"""
    N = int(N)
    if N <= 0:
        return 0
    if not contrib_flags:
        return NGRAD * N
    return ((N * 4 * NGRAD + 127) // 128 * 128 + (N + 127) // 128 * 128) // 4

c_fp = C.c_void_p  # all device pointers travel as void*


class TgsSettings(C.Structure):
    _fields_ = [
        ("image_width", C.c_int32), ("image_height", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("prefiltered", C.c_int32), ("debug", C.c_int32),
        ("tile_row_begin", C.c_int32), ("tile_row_end", C.c_int32),
        ("depth_normalize", C.c_int32), ("defer_count", C.c_int32), ("contrib_flags", C.c_int32),
        ("rendered_hint", C.c_int64),
        ("viewmatrix", c_fp), ("projmatrix", c_fp), ("campos", c_fp), ("bg", c_fp),
        ("alpha_max", C.c_float), ("near_z", C.c_float), ("principal_dx", C.c_float), ("principal_dy", C.c_float),
    ]


class TgsGaussians(C.Structure):
    _fields_ = [
        ("N", C.c_int32),
        ("means3D", c_fp), ("opacities", c_fp), ("shs", c_fp), ("colors_precomp", c_fp),
        ("scales", c_fp), ("rotations", c_fp), ("cov3D_precomp", c_fp),
    ]


class TgsTouch(C.Structure):
    _fields_ = [("target", c_fp), ("weight", c_fp), ("scale", c_fp), ("mode", C.c_int32),
                ("row_begin", C.c_int32), ("row_end", C.c_int32), ("grad_scale", c_fp)]


class TgsSaved(C.Structure):
    _fields_ = [("geom", c_fp), ("binning", c_fp), ("image", c_fp), ("num_rendered", C.c_int64),
                ("capacity", C.c_int64)]


class TgsGrads(C.Structure):
    _fields_ = [
        ("dmeans2D", c_fp), ("dmeans3D", c_fp), ("dopacity", c_fp), ("dshs", c_fp),
        ("dcolors", c_fp), ("dscales", c_fp), ("drotations", c_fp), ("dcov3D", c_fp),
    ]


class TgsGeomLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in
                ("records", "cov3D", "tiles_touched", "clamped", "rect", "depth_keys", "ids",
                 "depth_keys_sorted", "order", "span_sorted", "temp", "temp_bytes", "total")]


class TgsBinningLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in
                ("vals_sorted", "ckpt", "slot_tile", "ckpt_list", "work_counter", "slots", "total")]


class TgsImageLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("final_T", "n_contrib", "depth_raw", "color_acc", "ranges", "count", "total")]


class TgsRefBinningLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in
                ("offsets", "keys_unsorted", "keys_sorted", "vals_unsorted", "vals_sorted", "ranges", "temp",
                 "temp_bytes", "total")]


class TgsAdamGroup(C.Structure):
    _fields_ = [("param", c_fp), ("grad", c_fp), ("exp_avg", c_fp), ("exp_avg_sq", c_fp), ("numel", C.c_int64),
                ("lr", C.c_float), ("lr_tail", C.c_float), ("period", C.c_int32), ("head", C.c_int32)]


class TgsDensifyConfig(C.Structure):
    _fields_ = [("grad_thresh", C.c_float), ("size_thresh", C.c_float), ("cull_alpha_thresh", C.c_float),
                ("cull_scale_thresh", C.c_float), ("split_shrink", C.c_float), ("n_split_samples", C.c_int32),
                ("split_screen_radius", C.c_float), ("cull_screen_radius", C.c_float)]


class TgsParamSet(C.Structure):
    _fields_ = [("means", c_fp), ("shs", c_fp), ("opacity", c_fp), ("scales", c_fp), ("quats", c_fp)]


ADAM_MAX_GROUPS = 8

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_int, C.c_size_t)

# every symbol include/tgs.h declares: (restype, argtypes)
SIGNATURES = {
    "tgs_abi_version": (C.c_int, []),
    "tgs_last_error": (C.c_char_p, []),
    "tgs_launch_counts": (None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tgs_profile_enable": (C.c_int, [C.c_int32]),
    "tgs_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "tgs_forward_resolve": (C.c_int, [C.c_int64, C.c_int64, C.POINTER(C.c_int64)]),
    "tgs_mark_visible": (C.c_int, [C.c_int32, c_fp, c_fp, c_fp, c_fp]),
    "tgs_forward": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), ALLOC_FN, C.c_void_p,
                              c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, C.POINTER(TgsSaved), c_fp]),
    "tgs_screen_grad_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "tgs_backward_render": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), C.POINTER(TgsSaved),
                                      c_fp, c_fp, c_fp, C.POINTER(TgsTouch), c_fp, c_fp, c_fp]),
    "tgs_backward_preprocess": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), C.POINTER(TgsSaved),
                                          c_fp, c_fp, C.POINTER(TgsGrads), c_fp]),
    "tgs_backward_preprocess_gather": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), C.POINTER(TgsSaved),
                                                 c_fp, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32,
                                                 C.POINTER(TgsGrads), c_fp]),
    "tgs_backward": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), C.POINTER(TgsSaved), c_fp,
                               c_fp, c_fp, c_fp, C.POINTER(TgsTouch), c_fp, c_fp, C.POINTER(TgsGrads), c_fp]),
    "tgs_touch_loss_scale": (C.c_int, [c_fp, C.c_int64, C.c_float, C.c_float, c_fp, c_fp]),
    "tgs_touch_loss_value": (C.c_int, [c_fp, c_fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_fp, c_fp, c_fp,
                                       c_fp]),
    "tgs_fuse_touch_vision": (C.c_int, [c_fp, c_fp, c_fp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int32,
                                        C.c_double, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "tgs_decode_touch_maps": (C.c_int, [c_fp, c_fp, C.c_int64, C.c_float, C.c_float, C.c_int32, c_fp, c_fp, c_fp]),
    "tgs_train_step_host": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), c_fp, c_fp, c_fp,
                                      C.c_int32, C.c_float, C.POINTER(TgsGrads), c_fp, c_fp, c_fp, c_fp,
                                      C.POINTER(C.c_int64), c_fp]),
    "tgs_geom_layout": (C.c_int, [C.c_int32, C.POINTER(TgsGeomLayout)]),
    "tgs_binning_layout": (C.c_int, [C.c_int64, C.POINTER(TgsBinningLayout)]),
    "tgs_image_layout": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(TgsImageLayout)]),
    "tgs_photometric_scratch_floats": (C.c_size_t, [C.c_int32, C.c_int32]),
    "tgs_photometric_loss_forward": (C.c_int, [c_fp, c_fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                               c_fp, c_fp, c_fp, c_fp]),
    "tgs_photometric_loss_backward": (C.c_int, [c_fp, c_fp, c_fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                C.c_int32, C.c_int32, C.c_float, c_fp, c_fp, c_fp]),
    "tgs_activate_forward": (C.c_int, [C.c_int32, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "tgs_activate_backward": (C.c_int, [C.c_int32, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "tgs_adam_step": (C.c_int, [C.POINTER(TgsAdamGroup), C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, c_fp]),
    "tgs_densify_stats": (C.c_int, [C.c_int32, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "tgs_densify_temp_bytes": (C.c_size_t, [C.c_int32]),
    "tgs_densify_plan": (C.c_int, [C.c_int32, c_fp, c_fp, c_fp, c_fp, c_fp, C.POINTER(TgsDensifyConfig), C.c_int32,
                                   c_fp, c_fp, c_fp, C.c_size_t, C.POINTER(C.c_int64), c_fp]),
    "tgs_densify_apply": (C.c_int, [C.c_int32, C.c_int32, c_fp, c_fp, c_fp, C.POINTER(TgsDensifyConfig),
                                    C.POINTER(TgsParamSet), C.POINTER(TgsParamSet), c_fp, c_fp]),
    "tgs_project_gaussians": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), ALLOC_FN, C.c_void_p, c_fp,
                                        C.POINTER(TgsSaved), c_fp]),
    "tgs_project_gaussians_backward": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), C.POINTER(TgsSaved),
                                                 c_fp, c_fp, C.POINTER(TgsGrads), c_fp]),
    "tgs_rasterize_screen_forward": (C.c_int, [C.POINTER(TgsSettings), C.c_int32, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                               C.c_float, ALLOC_FN, C.c_void_p, c_fp, c_fp, c_fp, C.POINTER(TgsSaved), c_fp]),
    "tgs_rasterize_screen_backward": (C.c_int, [C.POINTER(TgsSettings), C.c_int32, C.POINTER(TgsSaved), c_fp, c_fp, c_fp,
                                                c_fp, c_fp]),
    "tgs_spherical_harmonics": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_fp, c_fp, c_fp, c_fp]),
    "tgs_spherical_harmonics_backward": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, c_fp, c_fp, c_fp, c_fp]),
    "tgs_refstructure_binning_layout": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.POINTER(TgsRefBinningLayout)]),
    "tgs_refstructure_forward": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians), ALLOC_FN, C.c_void_p,
                                           c_fp, c_fp, c_fp, c_fp, C.POINTER(TgsSaved), c_fp]),
    "tgs_refstructure_backward_render": (C.c_int, [C.POINTER(TgsSettings), C.POINTER(TgsGaussians),
                                                   C.POINTER(TgsSaved), c_fp, c_fp, c_fp, c_fp, c_fp]),
}

_lib = None


class TgsError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libtgs.so; raise loudly if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `python touch-gs_b200/build.py`). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale / a symbol is missing
        fn.restype = res
        fn.argtypes = args
    ver = lib.tgs_abi_version()
    if ver != TGS_ABI_VERSION:
        raise ImportError(f"libtgs.so ABI version {ver} != expected {TGS_ABI_VERSION}; rebuild")
    _lib = lib
    return lib


_ext = None
_ext_error = None


def load_ext():
    """The PyTorch C++ extension module `_C` (csrc/torch_ext.cpp -> _C.so next to libtgs.so): the operator's default
    binding (TORCH_CHECK argument errors, at::empty allocations, current CUDA stream, CUDAGuard).  Returns None when
    it has not been built; the caller then uses the ctypes binding of the same library (never a CPU path)."""
    global _ext, _ext_error
    if _ext is not None or _ext_error is not None:
        return _ext
    path = os.path.join(HERE, "_C.so")
    if os.environ.get("TGS_BINDING", "").lower() == "ctypes" or not os.path.exists(path):
        _ext_error = "disabled" if os.path.exists(path) else "not built"
        return None
    try:
        load()                                   # libtgs.so first (the extension links against it)
        import importlib.machinery
        import importlib.util
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        loader = importlib.machinery.ExtensionFileLoader("touchgs_b200_C", path)
        spec = importlib.util.spec_from_file_location("touchgs_b200_C", path, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        if mod.abi_version() != TGS_ABI_VERSION:
            raise ImportError(f"_C.so was built against ABI {mod.abi_version()}, expected {TGS_ABI_VERSION}; rebuild")
        _ext = mod
    except Exception as e:  # noqa: BLE001
        import warnings
        _ext_error = e
        warnings.warn(f"touchgs_b200: the torch extension _C.so failed to load ({type(e).__name__}: {e}); "
                      "using the ctypes binding of libtgs.so")
    return _ext


def use_binding(name: str) -> None:
    """Select the operator's binding for subsequent calls: "ext" (the torch C++ extension, default when built) or
    "ctypes".  For tests and the host-overhead measurement; both bindings drive the same kernels."""
    global _ext, _ext_error
    if name not in ("ext", "ctypes"):
        raise ValueError("binding must be 'ext' or 'ctypes'")
    os.environ["TGS_BINDING"] = "" if name == "ext" else "ctypes"
    _ext, _ext_error = None, None
    if name == "ext" and load_ext() is None:
        raise ImportError(f"the torch extension is unavailable: {_ext_error}")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tgs_last_error()
        raise TgsError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


STAGES = ("preprocess", "scan", "duplicate", "sort", "pack", "render_fwd", "loss_scale", "render_bwd",
          "preprocess_bwd", "photo_fwd", "photo_bwd", "activate", "adam", "refine", "bin_scatter")


def profile_enable(on: bool) -> None:
    load().tgs_profile_enable(1 if on else 0)


def profile_read():
    """{stage: (ms_total, launches)} since the last read; call after synchronising the stream."""
    ms = (C.c_float * len(STAGES))()
    cnt = (C.c_int32 * len(STAGES))()
    load().tgs_profile_read(ms, cnt)
    return {n: (float(ms[i]), int(cnt[i])) for i, n in enumerate(STAGES)}


def launch_counts():
    own, cub = C.c_uint64(0), C.c_uint64(0)
    load().tgs_launch_counts(C.byref(own), C.byref(cub))
    return int(own.value), int(cub.value)
