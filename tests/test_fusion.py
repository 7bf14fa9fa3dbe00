"""SURVEY §8(f) row N2 (touch / vision depth fusion -> touch target + weight): PINNED parity.

The golden vectors in tests/golden/fusion_reference.npz were produced by running the reference's own
functions (tests/golden/make_fusion_golden.py imports /root/reference/utils/fuse_touch_vision.py and
reads back the PNGs its save() wrote).  CPU tests: the numpy oracle against those bytes.  GPU tests: the
CUDA kernel (through the C ABI) against the golden bytes and, at full size, against the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import ROOT, T
from oracle import fusion_oracle as F

Z = np.load(os.path.join(ROOT, "tests", "golden", "fusion_reference.npz"))
OUT_KEYS = ("vision_aligned", "ds_gs", "fused", "fused_sigma")


def _case(i):
    sc = Z[f"case{i}_scalars"]
    ins = {k: Z[f"case{i}_in_{k}"] for k in ("touch", "vision", "touch_sigma")}
    outs = {k: Z[f"case{i}_out_{k}"] for k in OUT_KEYS}
    return ins, dict(scale=float(sc[0]), offset=float(sc[1]), offset2=float(sc[2]), is_real_world=bool(sc[3])), outs


@pytest.mark.parametrize("i", [0, 1])
def test_oracle_reproduces_reference_png_bytes(i):
    ins, sc, outs = _case(i)
    got = F.pipeline(ins["touch"], ins["vision"], ins["touch_sigma"], **sc)
    for k in OUT_KEYS:
        assert got[k].dtype == np.uint16 and np.array_equal(got[k], outs[k]), k
    # the vision-sigma restatement (zero-weight terms dropped) is exactly the reference's fp64 map
    _, v = F.align_apply(F.decode_mm(ins["vision"]), F.decode_mm(ins["touch"]), sc["scale"], sc["offset"],
                         sc["offset2"], sc["is_real_world"])
    assert np.array_equal(F.vision_sigma(v), Z[f"case{i}_vision_sigma_f64"])
    # touched pixels end up far more certain than the vision-only background (sigma ~5)
    touched = ins["touch_sigma"] > 0
    assert got["fused_sigma"][touched].max() < 100 and got["fused_sigma"][~touched].min() >= 5000


@pytest.mark.parametrize("i", [0, 1])
def test_device_math_on_host_reproduces_reference_png_bytes(host_math_lib, i):
    """csrc/touch_inputs_math.cuh (the kernel's per-pixel function) compiled for the host."""
    import ctypes
    ins, sc, outs = _case(i)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    t, v, s_ = (np.ascontiguousarray(ins[k]) for k in ("touch", "vision", "touch_sigma"))
    got = [np.zeros(t.shape, np.uint16) for _ in range(4)]
    tg, w = np.zeros(t.shape, np.float32), np.zeros(t.shape, np.float32)
    host_math_lib.hm_fuse(ctypes.c_long(t.size), P(t), P(v), P(s_), ctypes.c_double(sc["scale"]),
                          ctypes.c_double(sc["offset"]), ctypes.c_double(sc["offset2"]), int(sc["is_real_world"]),
                          ctypes.c_double(0.37), *[P(o) for o in got], P(tg), P(w))
    for k, g in zip(OUT_KEYS, got):
        assert np.array_equal(g, outs[k]), k
    rt, rw = F.training_decode(outs["fused"], outs["fused_sigma"], 0.37)
    assert np.array_equal(tg, rt) and np.array_equal(w, rw)


def test_training_decode_semantics():
    d = np.array([[0, 1500], [65535, 3]], np.uint16)
    s = np.array([[5000, 4], [0, 10000]], np.uint16)
    tgt, w = F.training_decode(d, s, scene_scale=0.5)
    assert tgt.dtype == np.float32 and w.dtype == np.float32
    np.testing.assert_allclose(tgt, [[0.0, 0.75], [32.7675, 0.0015]], rtol=1e-6)
    np.testing.assert_allclose(w, [[0.2, 250.0], [0.0, 0.1]], rtol=1e-6)


def _gpu_run(ins, sc, scene_scale=1.0):
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(v.copy()).to(dev) for k, v in ins.items()}
    out = T.touch_inputs.fuse_touch_vision(t["touch"], t["vision"], t["touch_sigma"], sc["scale"], sc["offset"],
                                           sc["offset2"], sc["is_real_world"], scene_scale)
    torch.cuda.synchronize()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("i", [0, 1])
def test_gpu_kernel_reproduces_reference_png_bytes(tgs_lib, i):
    ins, sc, outs = _case(i)
    got = _gpu_run(ins, sc, scene_scale=0.37)
    for k, g in zip(OUT_KEYS, got[:4]):
        assert np.array_equal(g.cpu().numpy(), outs[k]), k
    tgt, w = F.training_decode(outs["fused"], outs["fused_sigma"], 0.37)
    assert np.array_equal(got.target.cpu().numpy(), tgt) and np.array_equal(got.weight.cpu().numpy(), w)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(720, 1280), (37, 53), (1, 3), (900, 900)])
def test_gpu_kernel_matches_oracle_at_full_size(tgs_lib, shape):
    """Native image sizes of the reference data (1280x720 real, 900x900 sim: SURVEY A7) and ragged tails."""
    rng = np.random.default_rng(shape[0])
    h, w = shape
    touch = np.where(rng.random((h, w)) < 0.1, rng.integers(200, 3000, (h, w)), 0).astype(np.uint16)
    tsig = np.where(rng.random((h, w)) < 0.12, rng.integers(0, 60, (h, w)), 0).astype(np.uint16)
    vision = np.where(rng.random((h, w)) < 0.98, rng.integers(100, 6000, (h, w)), 0).astype(np.uint16)
    ins = dict(touch=touch, vision=vision, touch_sigma=tsig)
    for real in (True, False):
        sc = dict(scale=1.3127, offset=-0.2113, offset2=0.0171, is_real_world=real)
        ref = F.pipeline(touch, vision, tsig, scene_scale=1.9, **sc)
        got = _gpu_run(ins, sc, scene_scale=1.9)
        for k, g in zip(OUT_KEYS, got[:4]):
            assert np.array_equal(g.cpu().numpy(), ref[k]), (k, real)
        assert np.array_equal(got.target.cpu().numpy(), ref["target"])
        assert np.array_equal(got.weight.cpu().numpy(), ref["weight"])


@pytest.mark.gpu
def test_fused_touch_maps_feed_the_rasterizer(tgs_lib):
    """The kernel's fp32 outputs are exactly what GaussianRasterizer(touch_depth=, touch_weight=) takes."""
    ins, sc, _ = _case(0)
    got = _gpu_run(ins, sc)
    H, W = got.target.shape
    dev = got.target.device
    scn = T.synth.make_scene(300, 0, 0.05, 0.3, seed=3)
    cam = T.synth.look_at_camera(W, H, (0.2, 0.1, -3.0))
    rs = T.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0,
                                         cam.viewmatrix.to(dev), cam.projmatrix.to(dev), 0, cam.campos.to(dev))
    m = scn.means3D.to(dev).requires_grad_(True)
    color, _, depth, _, resid = T.GaussianRasterizer(rs)(
        m, None, scn.opacities.to(dev), shs=scn.shs.to(dev), scales=scn.scales.to(dev), rotations=scn.rotations.to(dev),
        touch_depth=got.target, touch_weight=got.weight, depth_loss="l1", depth_loss_mult=0.2)
    color.sum().backward()
    assert torch.isfinite(m.grad).all() and float(resid.abs().max()) > 0
