"""The PyTorch C++ extension `_C` (csrc/torch_ext.cpp; SURVEY §8b "C++/C-ABI surface"): exports, TORCH_CHECK errors,
and equivalence with the ctypes binding of the same library."""
import os
import time

import pytest
import torch

from helpers import T, synth, cuda_settings, rel_inf

L = T._lib


@pytest.fixture(scope="module")
def ext(tgs_lib):
    import importlib
    importlib.import_module("touch-gs_b200.build").build_torch_ext()
    L.use_binding("ext")
    e = L.load_ext()
    assert e is not None
    return e


def test_extension_exports_the_reference_era_entry_points(ext):
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible", "backward_render",
                 "backward_preprocess", "touch_loss_scale", "touch_loss_value"):
        assert callable(getattr(ext, name)), name
    assert ext.abi_version() == L.TGS_ABI_VERSION
    assert any("_C.so" in l for l in open("/proc/self/maps"))


def test_torch_check_errors_without_gpu(ext):
    z = torch.zeros
    e = z(0)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ext.mark_visible(z(4, 3), torch.eye(4), torch.eye(4))
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        ext.mark_visible(z(4, 2), torch.eye(4), torch.eye(4))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ext.rasterize_gaussians(z(3), z(4, 3), e, z(4), z(4, 3), z(4, 4), 1.0, e, torch.eye(4), torch.eye(4), 0.5, 0.5, 64, 64,
                                z(4, 1, 3), 0, z(3), False, False, 0, 4, True, 0, None, False)


@pytest.mark.gpu
def test_torch_check_errors_on_device(ext):
    dev = torch.device("cuda:0")
    z = lambda *s, **k: torch.zeros(*s, device=dev, **k)
    e = z(0)
    eye = torch.eye(4, device=dev)
    base = dict(bg=z(3), means=z(4, 3), col=e, op=z(4), sc=z(4, 3), rot=z(4, 4), cov=e, sh=z(4, 1, 3))

    def call(**kw):
        a = dict(base, **kw)
        return ext.rasterize_gaussians(a["bg"], a["means"], a["col"], a["op"], a["sc"], a["rot"], 1.0, a["cov"], eye, eye, 0.5, 0.5,
                                       64, 64, a["sh"], 0, z(3), False, False, 0, 4, True, 0, None, False)
    call()                                                        # the valid call goes through
    with pytest.raises(RuntimeError, match="exactly one of either SHs or precomputed colors"):
        call(col=z(4, 3))
    with pytest.raises(RuntimeError, match="scale/rotation pair or precomputed 3D covariance"):
        call(cov=z(4, 6))
    with pytest.raises(RuntimeError, match="scales must have shape"):
        call(sc=z(5, 3))
    with pytest.raises(RuntimeError, match="rotations must be float32"):
        call(rot=z(4, 4, dtype=torch.float64))
    with pytest.raises(RuntimeError, match="viewmatrix must be on"):
        ext.rasterize_gaussians(z(3), z(4, 3), e, z(4), z(4, 3), z(4, 4), 1.0, e, torch.eye(4), eye, 0.5, 0.5, 64, 64, z(4, 1, 3), 0,
                                z(3), False, False, 0, 4, True, 0, None, False)
    with pytest.raises(RuntimeError, match="tile rows"):
        ext.rasterize_gaussians(z(3), z(4, 3), e, z(4), z(4, 3), z(4, 4), 1.0, e, eye, eye, 0.5, 0.5, 64, 64, z(4, 1, 3), 0, z(3),
                                False, False, 3, 9, True, 0, None, False)


@pytest.mark.gpu
def test_extension_and_ctypes_bindings_agree_and_host_overhead(ext):
    dev = torch.device("cuda:0")
    sc = synth.make_scene(3000, 2, 0.02, 0.3, seed=2)
    cam = synth.look_at_camera(203, 117, (0.2, -0.4, -2.2))
    rs = cuda_settings(cam, 2, dev, (0.1, 0.2, 0.3))
    g = torch.Generator().manual_seed(0)
    grgb = (torch.rand(3, 117, 203, generator=g) / (3 * 117 * 203)).to(dev)
    with torch.no_grad():
        d0 = T.GaussianRasterizer(rs)(sc.means3D.to(dev), None, sc.opacities.to(dev), shs=sc.shs.to(dev), scales=sc.scales.to(dev),
                                      rotations=sc.rotations.to(dev))[2]
    tgt, wgt = synth.make_touch_maps(d0[0].cpu() + 0.02, seed=1, n_patches=3, patch_radius=10)
    tgt, wgt = tgt.to(dev), wgt.to(dev)
    names = ("means3D", "scales", "rotations", "opacities", "shs")

    def run(binding, steps=1):
        L.use_binding(binding)
        ins = {k: getattr(sc, k).to(dev).clone().requires_grad_(True) for k in names}
        m2d = torch.zeros(3000, 3, device=dev, requires_grad=True)
        out = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            for v in ins.values():
                v.grad = None
            out = T.GaussianRasterizer(rs)(ins["means3D"], m2d, ins["opacities"], shs=ins["shs"], scales=ins["scales"],
                                           rotations=ins["rotations"], touch_depth=tgt, touch_weight=wgt, depth_loss="l1",
                                           depth_loss_mult=0.2, return_touch_loss=True, tile_rows=(1, 7))
            ((out[0] * grgb).sum() + out[5]).backward()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        return out, {k: v.grad.clone() for k, v in ins.items()}, m2d.grad.clone(), dt
    try:
        oe, ge, me, _ = run("ext")
        oc, gc, mc, _ = run("ctypes")
        for a, b in zip(oe, oc):
            assert torch.equal(a, b), "forward outputs of the two bindings differ"
        for k in names:
            assert rel_inf(ge[k], gc[k]) < 1e-5, k                 # same kernels; atomics reorder the sums
        assert rel_inf(me, mc) < 1e-5
        # host overhead of a (tiny, launch-bound) forward + backward through each binding
        run("ext", 20); run("ctypes", 20)
        te = min(run("ext", 200)[3] for _ in range(3))
        tc = min(run("ctypes", 200)[3] for _ in range(3))
        print(f"\\nhost-bound step (3000 Gaussians, 203x117, fwd+bwd): ext {te * 1e6:.0f} us, ctypes {tc * 1e6:.0f} us")
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "host_overhead.txt"), "w") as f:
            f.write(f"launch-bound operator step (3000 Gaussians, 203x117, forward + fused-touch backward, 200 steps, best of 3):\\n"
                    f"torch C++ extension _C: {te * 1e6:.1f} us/step\\nctypes binding:        {tc * 1e6:.1f} us/step\\n")
    finally:
        L.use_binding("ext")
