"""CPU checks of the drop-in boundary: libtgs.so loads, exports every symbol include/tgs.h declares,
its struct layouts match the ctypes mirror, and argument errors are reported without touching a GPU."""
import ctypes as C
import os
import re

import pytest

from helpers import ROOT, T


def _declared_functions():
    txt = open(os.path.join(ROOT, "include", "tgs.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b(tgs_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(n for n in names if n != "tgs_alloc_fn"))


def test_header_symbols_exported(tgs_lib):
    decl = _declared_functions()
    assert len(decl) >= 12
    for name in decl:
        assert hasattr(tgs_lib, name), f"{name} declared in include/tgs.h but not exported by libtgs.so"
    assert set(decl) == set(T._lib.SIGNATURES), "ctypes SIGNATURES out of sync with include/tgs.h"
    assert tgs_lib.tgs_abi_version() == T._lib.TGS_ABI_VERSION


def test_layouts_are_aligned_and_monotone(tgs_lib):
    L = T._lib
    g = L.TgsGeomLayout(); tgs_lib.tgs_geom_layout(1000, C.byref(g))
    assert g.records == 0 and g.cov3D >= 48 * 1000 and g.total >= g.rect + 8 * 1000
    for n, _ in g._fields_:
        assert getattr(g, n) % 256 == 0
    b = L.TgsBinningLayout(); tgs_lib.tgs_binning_layout(5000, C.byref(b))
    assert b.vals_sorted == 0 and b.ckpt >= b.vals_sorted + 4 * 5000           # no per-instance records: ids only
    assert b.slots == (5000 >> 8) + 2 and b.slot_tile >= b.ckpt + b.slots * 5 * 256 * 4 and b.total % 256 == 0
    i = L.TgsImageLayout(); tgs_lib.tgs_image_layout(100, 50, C.byref(i))
    assert i.total >= 6 * 4 * 5000 and i.count >= i.ranges + 7 * 4 * 8          # 7 x 4 tiles of (start, end)


def test_screen_grad_buffer_size_and_settings_layout(tgs_lib):
    """contrib_flags sits in what used to be alignment padding of TgsSettings; the Python size helper mirrors the C one."""
    L = T._lib
    assert L.TgsSettings.defer_count.offset == 48 and L.TgsSettings.contrib_flags.offset == 52
    assert L.TgsSettings.rendered_hint.offset == 56
    for n in (0, 1, 3, 127, 128, 129, 30000, 1_000_000, 5_000_001):
        for f in (0, 1):
            b = tgs_lib.tgs_screen_grad_bytes(n, f)
            assert b == 4 * L.screen_grad_floats(n, bool(f)), (n, f)
            if n and f:
                off = (n * 40 + 127) // 128 * 128
                assert b >= off + n and off % 128 == 0 and b % 128 == 0
            elif n:
                assert b == 40 * n


def test_argument_errors_without_gpu(tgs_lib):
    L = T._lib
    s = L.TgsSettings(image_width=64, image_height=64, tanfovx=0.5, tanfovy=0.5, scale_modifier=1.0)
    g = L.TgsGaussians(N=4)
    saved = L.TgsSaved()
    cb = L.ALLOC_FN(lambda u, w, n: None)
    rc = tgs_lib.tgs_forward(C.byref(s), C.byref(g), cb, None, None, None, None, None, None, None, C.byref(saved), None)
    assert rc == -1 and b"viewmatrix" in tgs_lib.tgs_last_error()
    s.viewmatrix = s.projmatrix = s.bg = s.campos = 0x1000      # never dereferenced: validation fails first
    g.means3D = g.opacities = 0x1000
    rc = tgs_lib.tgs_forward(C.byref(s), C.byref(g), cb, None, None, None, None, None, None, None, C.byref(saved), None)
    assert rc == -1 and b"exactly one of either SHs or precomputed colors" in tgs_lib.tgs_last_error()
    g.colors_precomp = 0x1000
    g.scales = 0x1000
    rc = tgs_lib.tgs_forward(C.byref(s), C.byref(g), cb, None, None, None, None, None, None, None, C.byref(saved), None)
    assert rc == -1 and b"scale/rotation pair" in tgs_lib.tgs_last_error()
    rc = tgs_lib.tgs_mark_visible(3, None, None, None, None)
    assert rc == -1


def test_operator_rejects_cpu_tensors_and_bad_combos(tgs_lib):
    import torch
    cam = T.synth.look_at_camera(32, 32, (0.0, 0.0, -3.0))
    rs = T.GaussianRasterizationSettings(32, 32, cam.tanfovx, cam.tanfovy, torch.zeros(3), 1.0, cam.viewmatrix,
                                         cam.projmatrix, 0, cam.campos)
    r = T.GaussianRasterizer(rs)
    m = torch.zeros(4, 3); o = torch.ones(4, 1); sh = torch.zeros(4, 1, 3); sc = torch.ones(4, 3); q = torch.zeros(4, 4)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(m, None, o, shs=None, colors_precomp=None, scales=sc, rotations=q)
    with pytest.raises(Exception, match="scale/rotation pair"):
        r(m, None, o, shs=sh, scales=sc, rotations=q, cov3D_precomp=torch.zeros(4, 6))
    with pytest.raises(RuntimeError, match="CUDA-only"):      # loud failure, never a CPU fallback
        r(m, None, o, shs=sh, scales=sc, rotations=q)


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under touch-gs_b200/ may reference it."""
    pkg = os.path.join(ROOT, "touch-gs_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import oracle|from oracle)", txt, flags=re.M), os.path.join(dp, f)


def test_scratch_allocator_has_no_reference_cycle(tgs_lib):
    """The ctypes allocator callback must not keep the ~1 GB saved buffers alive until the cyclic GC
    runs (regression: a bound-method callback made memory_reserved grow to tens of GB)."""
    import gc
    import weakref
    import torch
    from importlib import import_module
    R = import_module("touch-gs_b200.rasterizer")
    gc.disable()
    try:
        s = R._Scratch(torch.device("cpu"))
        ptr = s.cb(None, 1, 1024)
        assert ptr and 1 in s.bufs
        w = weakref.ref(s.bufs[1])
        del s
        assert w() is None, "scratch buffer survived refcount release: reference cycle"
    finally:
        gc.enable()


def test_new_entry_points_validate_arguments_without_gpu(tgs_lib):
    """Train-step, screen-space and multi-GPU entry points reject bad arguments before touching a device."""
    L = T._lib
    lib = tgs_lib
    assert lib.tgs_photometric_loss_forward(None, None, 8, 8, 0, 0, 0.2, None, None, None, None) == -1
    assert b"tgs_photometric_loss_forward" in lib.tgs_last_error()
    assert lib.tgs_photometric_loss_backward(None, None, None, 8, 8, 0, 0, 0, 0, 0.2, None, None, None) == -1
    assert lib.tgs_photometric_scratch_floats(10, 7) == 9 * 70
    assert lib.tgs_activate_forward(4, None, None, None, None, None, None, None) == -1
    assert lib.tgs_activate_forward(0, None, None, None, None, None, None, None) == 0            # empty input is fine
    g = (L.TgsAdamGroup * 1)(L.TgsAdamGroup(param=0x1000, grad=0x1000, exp_avg=0x1000, exp_avg_sq=0x1004, numel=16, lr=1e-3,
                                            lr_tail=1e-3, period=0, head=0))
    assert lib.tgs_adam_step(g, 1, 1, 0.9, 0.999, 1e-15, None) == -1 and b"16-byte aligned" in lib.tgs_last_error()
    assert lib.tgs_adam_step(g, 0, 1, 0.9, 0.999, 1e-15, None) == -1
    assert lib.tgs_adam_step(g, 1, 0, 0.9, 0.999, 1e-15, None) == -1                            # step counts from 1
    g[0].exp_avg_sq = 0x1000
    g[0].period, g[0].head = 48, 49
    assert lib.tgs_adam_step(g, 1, 1, 0.9, 0.999, 1e-15, None) == -1 and b"period/head" in lib.tgs_last_error()
    assert lib.tgs_densify_stats(5, None, None, None, None, None, None) == -1
    cfg = L.TgsDensifyConfig(grad_thresh=2e-4, size_thresh=0.01, cull_alpha_thresh=0.1, cull_scale_thresh=0.5,
                             split_shrink=1.6, n_split_samples=99)
    tot = C.c_int64(0)
    tot.value = 7
    assert lib.tgs_densify_plan(0, None, None, None, None, None, C.byref(cfg), 1, None, None, None, 0, C.byref(tot), None) == 0
    assert tot.value == 0                                                                      # empty population stays empty
    assert lib.tgs_densify_plan(-1, None, None, None, None, None, C.byref(cfg), 1, None, None, None, 0, C.byref(tot), None) == -1
    assert lib.tgs_densify_plan(8, None, None, None, None, None, C.byref(cfg), 1, None, None, None, 0, C.byref(tot), None) == -1
    assert lib.tgs_densify_plan(8, 0x1000, 0x1000, 0x1000, 0x1000, None, C.byref(cfg), 1, 0x1000, 0x1000, 0x1000, 1024,
                                C.byref(tot), None) == -1 and b"n_split_samples" in lib.tgs_last_error()
    assert lib.tgs_touch_loss_value(None, None, 8, 8, 0, 0, 1, None, None, None, None) == -1
    assert lib.tgs_touch_loss_value(0x1000, None, 8, 8, 0, 0, 7, 0x1000, 0x1000, 0x1000, None) == -1 and b"bad mode" in lib.tgs_last_error()
    assert lib.tgs_densify_temp_bytes(1000) >= 256
    s = L.TgsSettings(image_width=64, image_height=64, tanfovx=0.5, tanfovy=0.5, scale_modifier=1.0)
    gg = L.TgsGaussians(N=4)
    saved = L.TgsSaved()
    cb = L.ALLOC_FN(lambda u, w, n: None)
    assert lib.tgs_project_gaussians(C.byref(s), C.byref(gg), cb, None, None, C.byref(saved), None) == -1
    assert b"viewmatrix" in lib.tgs_last_error()
    assert lib.tgs_rasterize_screen_forward(C.byref(s), 4, None, None, None, None, None, None, 0.5, cb, None, None, None, None,
                                            C.byref(saved), None) == -1
    assert lib.tgs_spherical_harmonics(4, 5, 16, None, None, None, None) == -1                  # degree > 3
    assert lib.tgs_spherical_harmonics(4, 3, 9, 0x1000, 0x1000, 0x1000, None) == -1             # K < (deg+1)^2
    s.viewmatrix = s.projmatrix = s.bg = s.campos = 0x1000
    gg.means3D = gg.opacities = gg.colors_precomp = gg.scales = gg.rotations = 0x1000
    saved.geom = 0x1000
    ptrs = (C.c_void_p * 2)(0x1000, 0)
    rows = (C.c_int32 * 4)(0, 2, 2, 4)
    gr = L.TgsGrads()
    assert lib.tgs_backward_preprocess_gather(C.byref(s), C.byref(gg), C.byref(saved), 0x1000, ptrs, rows, 0, C.byref(gr), None) == -1
    assert lib.tgs_backward_preprocess_gather(C.byref(s), C.byref(gg), C.byref(saved), 0x1000, ptrs, rows, 9, C.byref(gr), None) == -1
    assert lib.tgs_backward_preprocess_gather(C.byref(s), C.byref(gg), C.byref(saved), 0x1000, ptrs, rows, 2, C.byref(gr), None) == -1
    assert b"peer 1 pointer is NULL" in lib.tgs_last_error()
    lay = L.TgsRefBinningLayout()
    assert lib.tgs_refstructure_binning_layout(100, 1000, 64, C.byref(lay)) == 0
    assert lay.keys_sorted >= lay.keys_unsorted + 8000 and lay.total % 256 == 0


def test_struct_sizes_match_header(tgs_lib):
    """ctypes mirrors vs the C structs: compile a tiny C program against include/tgs.h and compare sizeof()."""
    import subprocess
    import tempfile
    L = T._lib
    names = ["TgsSettings", "TgsGaussians", "TgsTouch", "TgsSaved", "TgsGrads", "TgsGeomLayout", "TgsBinningLayout",
             "TgsImageLayout", "TgsRefBinningLayout", "TgsAdamGroup", "TgsDensifyConfig", "TgsParamSet"]
    src = '#include <stdio.h>\n#include "tgs.h"\nint main(void){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for n in names:
        assert int(sizes[n]) == C.sizeof(getattr(L, n)), f"{n}: C {sizes[n]} != ctypes {C.sizeof(getattr(L, n))}"


def test_gsplat_style_surface_rejects_cpu_tensors(tgs_lib):
    import torch
    from importlib import import_module
    G = import_module("touch-gs_b200.gsplat_compat")
    z = torch.zeros
    with pytest.raises(RuntimeError, match="CUDA-only"):
        G.project_gaussians(z(4, 3), z(4, 3), 1.0, torch.ones(4, 4), torch.eye(4), torch.eye(4), 100.0, 100.0, 32.0, 32.0, 64, 64)
    with pytest.raises(ValueError, match="C <= 3"):
        G.rasterize_gaussians(z(4, 2), z(4), z(4, dtype=torch.int32), z(4, 3), z(4, dtype=torch.int32), z(4, 5), z(4, 1), 64, 64)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        G.spherical_harmonics(3, z(4, 3), z(4, 16, 3))
    assert G.PIXEL_CENTER_OFFSET == 0.5 and G.ALPHA_MAX == 0.999


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent fallback: without the built CUDA library the product refuses to load (and says how to build it)."""
    L = T._lib
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "libtgs.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        L.load()
