"""CPU check of the per-Gaussian device math (csrc/tgs_math.cuh compiled for the host, FMA
contraction off) against the oracle: forward BIT-EXACT on everything that feeds integers, backward
(our own chain-rule derivation) against oracle autograd."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import O, synth, oracle_settings


class Cam(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("fx", "fy", "limx", "limy", "mod")] + \
               [(n, ctypes.c_int) for n in ("W", "H", "Tx", "Ty", "row0", "row1", "deg", "K")] + \
               [(n, ctypes.c_float) for n in ("near_z", "ppx", "ppy", "alpha_max")]


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


CASES = [
    dict(N=2000, W=128, H=128, deg=3, smin=0.02, smax=0.2, eye=(0.5, 0.3, -3.0), seed=0),
    dict(N=5000, W=320, H=200, deg=2, smin=0.004, smax=0.04, eye=(2.0, 1.0, -2.0), seed=1, mod=1.3),
    dict(N=3000, W=200, H=120, deg=1, smin=0.05, smax=0.5, eye=(0.2, 0.1, -1.2), seed=2, band=(2, 5)),
    dict(N=3000, W=64, H=64, deg=0, smin=0.2, smax=1.5, eye=(0.0, 0.0, -0.9), seed=3),
    dict(N=1000, W=1920, H=1080, deg=3, smin=0.002, smax=0.02, eye=(0.0, 0.5, -3.0), seed=4),
]


@pytest.mark.parametrize("case", CASES)
def test_forward_bit_exact_and_backward(host_math_lib, case):
    lib = host_math_lib
    N, W, H, deg = case["N"], case["W"], case["H"], case["deg"]
    mod, band = case.get("mod", 1.0), case.get("band")
    sc = synth.make_scene(N, deg, case["smin"], case["smax"], seed=case["seed"])
    cam = synth.look_at_camera(W, H, case["eye"])
    S = oracle_settings(cam, deg, mod=mod)
    fx, fy, lx, ly = O.camera_scalars(S)
    Tx, Ty = (W + 15) // 16, (H + 15) // 16
    b = band or (0, Ty)
    K = (deg + 1) ** 2
    c = Cam(fx, fy, lx, ly, mod, W, H, Tx, Ty, b[0], b[1], deg, K, 0.2, 0.0, 0.0, 0.99)
    ins = [t.clone().requires_grad_(True) for t in (sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)]
    m, s, r, o, sh = ins
    pre = O.preprocess(m, s, r, o, sh, None, None, S, band)
    f = np.zeros((N, 6), np.float32)
    ii = np.zeros((N, 6), np.int32)
    rgb = np.zeros((N, 3), np.float32)
    cl = np.zeros(N, np.uint32)
    cov = np.zeros((N, 6), np.float32)
    A = [x.detach().numpy() for x in (m, s, r, sh)]
    vm, pm, cp = cam.viewmatrix.numpy().copy(), cam.projmatrix.numpy().copy(), cam.campos.numpy().copy()
    lib.hm_preprocess(N, P(A[0]), P(A[1]), P(A[2]), P(A[3]), None, P(vm), P(pm), P(cp), ctypes.byref(c),
                      P(f), P(ii), P(rgb), P(cl), P(cov))
    vis = pre.radii.numpy() > 0
    assert vis.sum() > 0
    assert (ii[:, 0] == pre.radii.numpy()).all()
    assert (ii[:, 5] == pre.tiles_touched.numpy()).all()
    rm = np.concatenate([pre.rect_min.numpy(), pre.rect_max.numpy()], 1)
    assert (ii[vis, 1:5] == rm[vis]).all()
    assert (f[vis, 0:2].view(np.int32) == pre.xy.detach().numpy()[vis].view(np.int32)).all()
    assert (f[vis, 2].view(np.int32) == pre.depth.detach().numpy()[vis].view(np.int32)).all()
    assert (f[vis, 3:6].view(np.int32) == pre.conic.detach().numpy()[vis].view(np.int32)).all()
    assert (cov.view(np.int32) == pre.cov3D.detach().numpy().view(np.int32)).all()
    assert np.abs(rgb[vis] - pre.rgb.detach().numpy()[vis]).max() < 1e-5
    assert (((cl[vis][:, None] >> np.arange(3)) & 1) == pre.clamped.numpy()[vis]).all()

    g = torch.Generator().manual_seed(case["seed"] + 5)
    sg = torch.randn(N, 10, generator=g) * torch.tensor(vis)[:, None]
    L = ((pre.xy * sg[:, 0:2]).sum() + (pre.conic * sg[:, 2:5]).sum() + (pre.opacity * sg[:, 5]).sum()
         + (pre.rgb * sg[:, 6:9]).sum() + (pre.depth * sg[:, 9]).sum())
    L.backward()
    dm = np.zeros((N, 3), np.float32)
    ds = np.zeros((N, 3), np.float32)
    dq = np.zeros((N, 4), np.float32)
    dsh = np.zeros((N, K, 3), np.float32)
    dcv = np.zeros((N, 6), np.float32)
    sgn, rad = sg.numpy().copy(), pre.radii.numpy().copy()
    lib.hm_backward(N, P(A[0]), P(A[1]), P(A[2]), P(A[3]), None, P(vm), P(pm), P(cp), ctypes.byref(c), P(rad),
                    P(cl), P(sgn), P(dm), P(ds), P(dq), P(dsh), P(dcv))
    for name, a, ref in (("means", dm, m.grad), ("scales", ds, s.grad), ("rots", dq, r.grad), ("sh", dsh, sh.grad)):
        ref = ref.numpy()
        err = np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-20)
        assert err < 1e-4, (name, err)
        # element-wise, per Gaussian scale (gradient magnitudes span many decades across Gaussians)
        den = np.abs(ref).reshape(N, -1).max(1).reshape((N,) + (1,) * (ref.ndim - 1)) + 1e-30
        assert (np.abs(a - ref) / den).max() < 2e-3, name


def test_precomputed_cov3d_path(host_math_lib):
    lib = host_math_lib
    N, W, H, deg = 500, 96, 64, 0
    sc = synth.make_scene(N, deg, 0.03, 0.3, seed=9)
    cam = synth.look_at_camera(W, H, (0.3, 0.2, -3.0))
    S = oracle_settings(cam, deg)
    pre0 = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, None, None, S)
    cov_in = pre0.cov3D.clone().requires_grad_(True)
    m = sc.means3D.clone().requires_grad_(True)
    pre = O.preprocess(m, None, None, sc.opacities, sc.shs, None, cov_in, S)
    assert torch.equal(pre.radii, pre0.radii)
    fx, fy, lx, ly = O.camera_scalars(S)
    c = Cam(fx, fy, lx, ly, 1.0, W, H, 6, 4, 0, 4, deg, 1, 0.2, 0.0, 0.0, 0.99)
    g = torch.Generator().manual_seed(1)
    vis = pre.radii.numpy() > 0
    sg = torch.randn(N, 10, generator=g) * torch.tensor(vis)[:, None]
    ((pre.xy * sg[:, 0:2]).sum() + (pre.conic * sg[:, 2:5]).sum() + (pre.depth * sg[:, 9]).sum()).backward()
    dm = np.zeros((N, 3), np.float32); ds = np.zeros((N, 3), np.float32); dq = np.zeros((N, 4), np.float32)
    dcv = np.zeros((N, 6), np.float32)
    A = [sc.means3D.numpy(), cov_in.detach().numpy()]
    vm, pm, cp = cam.viewmatrix.numpy().copy(), cam.projmatrix.numpy().copy(), cam.campos.numpy().copy()
    cl = np.zeros(N, np.uint32); sgn = sg.numpy().copy(); rad = pre.radii.numpy().copy()
    lib.hm_backward(N, P(A[0]), None, None, None, P(A[1]), P(vm), P(pm), P(cp), ctypes.byref(c), P(rad), P(cl),
                    P(sgn), P(dm), P(ds), P(dq), None, P(dcv))
    for a, ref in ((dm, m.grad.numpy()), (dcv, cov_in.grad.numpy())):
        assert np.abs(a - ref).max() / np.abs(ref).max() < 1e-4
