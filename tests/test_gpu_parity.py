"""GPU parity tests: the CUDA path (through the C ABI of libtgs.so) against the CPU oracle on the
same seeded inputs.

Bars (north star): BIT-EXACT on tile / sort indices (radii, rects, tiles_touched, scan offsets,
unsorted and sorted (key, value) arrays, tile ranges); 1e-4 relative on RGB / depth / alpha /
gradient tensors.  Two documented caveats: exp() on the GPU (ex2.approx) and on the CPU differ in the
last ulps, so a (pixel, splat) pair sitting exactly on the alpha >= 1/255 or T < 1e-4 threshold can
flip -- n_contrib is therefore compared with a tiny mismatch budget, and image-like tensors may have
a few-pixel outlier budget (see helpers.assert_close_tensor).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from helpers import O, T, synth, oracle_settings, cuda_settings, rel_inf, assert_close_tensor, ROOT

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")


@pytest.fixture(scope="module", autouse=True)
def _require_cuda(tgs_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    own0, _ = T._lib.launch_counts()
    yield
    own1, _ = T._lib.launch_counts()
    assert own1 > own0, "no libtgs kernels were launched: native path not exercised"
    assert any("libtgs.so" in l for l in open("/proc/self/maps")), "libtgs.so not loaded"


def _to(dev, *ts):
    return [None if t is None else t.to(dev) for t in ts]


CASES = {
    "c1": dict(N=1000, W=128, H=128, deg=0, smin=0.02, smax=0.2, eye=(0.5, 0.3, -3.0), seed=0),
    "deg3": dict(N=20000, W=320, H=200, deg=3, smin=0.006, smax=0.06, eye=(1.5, 0.8, -2.5), seed=1),
    "ragged": dict(N=3000, W=203, H=117, deg=2, smin=0.02, smax=0.3, eye=(0.2, -0.4, -2.2), seed=2, mod=1.2),
    "close": dict(N=1500, W=96, H=80, deg=1, smin=0.1, smax=0.8, eye=(0.0, 0.1, -1.1), seed=3),
    "band": dict(N=4000, W=160, H=160, deg=1, smin=0.02, smax=0.2, eye=(0.4, 0.2, -3.0), seed=4, band=(3, 7)),
    "band_sh3": dict(N=6000, W=192, H=176, deg=3, smin=0.01, smax=0.15, eye=(0.3, 0.2, -2.8), seed=5, band=(2, 6)),
}


def _case(name):
    c = CASES[name]
    sc = synth.make_scene(c["N"], c["deg"], c["smin"], c["smax"], seed=c["seed"])
    cam = synth.look_at_camera(c["W"], c["H"], c["eye"])
    return c, sc, cam


@pytest.mark.parametrize("name", list(CASES))
def test_preprocess_and_binning_bit_exact(name):
    c, sc, cam = _case(name)
    band, mod = c.get("band"), c.get("mod", 1.0)
    S = oracle_settings(cam, c["deg"], mod=mod)
    pre = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, None, None, S, band)
    bins = O.bin_and_sort(pre, S)
    rs = cuda_settings(cam, c["deg"], DEV, mod=mod, debug=True)
    m, s, r, o, sh = _to(DEV, sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)
    st = T.inspect_state.forward_state(m, o, rs, shs=sh, scales=s, rotations=r, opt=T.TouchOptions(tile_rows=band))
    vis = pre.radii > 0
    assert int(vis.sum()) > 0
    assert torch.equal(st["radii"].cpu(), pre.radii)
    assert torch.equal(st["tiles_touched"].cpu(), pre.tiles_touched)
    assert torch.equal(st["rect_min"].cpu()[vis], pre.rect_min[vis])
    assert torch.equal(st["rect_max"].cpu()[vis], pre.rect_max[vis])
    assert st["num_rendered"] == bins.keys.numel() == st["num_rendered_device"]
    # depth order of the Gaussians (phase 1 of the two-phase sort): ascending (depth bits, id) over emitters
    dk = torch.where(pre.tiles_touched > 0, pre.depth.detach().contiguous().view(torch.int32).long() & 0xFFFFFFFF,
                     torch.full_like(pre.radii, 0xFFFFFFFF, dtype=torch.int64))
    assert torch.equal(st["order"].cpu().long(), torch.sort(dk, stable=True).indices)
    bits = lambda t: t.contiguous().view(torch.int32)
    assert torch.equal(bits(st["xy"].cpu()[vis]), bits(pre.xy[vis]))
    assert torch.equal(bits(st["gdepth"].cpu()[vis]), bits(pre.depth[vis]))
    assert torch.equal(bits(st["conic"].cpu()[vis]), bits(pre.conic[vis]))
    assert torch.equal(bits(st["cov3D"].cpu()), bits(pre.cov3D))
    # no per-instance keys are ever emitted or sorted (binning.cu counts rectangles): the final sorted list and the
    # tile ranges are what the spec pins; as a multiset the list must equal the oracle's Gaussian-major emission
    listed = (st["tile_ids"].cpu() << 32) | st["vals"].cpu().long()
    ref_emitted = ((bins.keys_unsorted >> 32) << 32) | bins.vals_unsorted.long()
    assert torch.equal(torch.sort(listed).values, torch.sort(ref_emitted).values)
    assert torch.equal(st["keys"].cpu(), bins.keys)
    assert torch.equal(st["vals"].cpu(), bins.vals)
    assert torch.equal(st["ranges"].cpu(), bins.ranges)
    # colour (and its clamp mask) is evaluated only for the Gaussians that emit instances on this rank (inside the band);
    # the other visible ones carry the mask "unknown" (0x80) when the SH rows are staged (K = 16), see preprocess.cu
    em = vis & (pre.tiles_touched > 0)
    assert float((st["rgb"].cpu()[em] - pre.rgb[em]).abs().max()) < 1e-5
    cl = st["clamped"].cpu()[em]
    for ch in range(3):
        assert torch.equal(((cl >> ch) & 1).bool(), pre.clamped[em][:, ch])
    rest = st["clamped"].cpu()[vis & ~em]
    if rest.numel():
        staged = int(sc.shs.shape[1]) == 16
        assert bool(((rest & 0x80) != 0).all()) if staged else True
    # packed records = per-Gaussian records gathered in sorted order
    assert torch.equal(st["records"].cpu()[:, 3].contiguous().view(torch.int32), bins.vals)


@pytest.mark.parametrize("name", list(CASES))
def test_render_forward(name):
    c, sc, cam = _case(name)
    band, mod = c.get("band"), c.get("mod", 1.0)
    bg = (0.3, 0.1, 0.2)
    S = oracle_settings(cam, c["deg"], bg, mod)
    ref = O.rasterize(sc.means3D, sc.opacities, S, shs=sc.shs, scales=sc.scales, rotations=sc.rotations, band=band)
    rs = cuda_settings(cam, c["deg"], DEV, bg, mod)
    m, s, r, o, sh = _to(DEV, sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)
    st = T.inspect_state.forward_state(m, o, rs, shs=sh, scales=s, rotations=r, opt=T.TouchOptions(tile_rows=band))
    H, W = c["H"], c["W"]
    y0, y1 = (0, H) if band is None else T.sharding.band_pixel_rows(band, H)
    sl = (slice(None), slice(y0, y1))
    budget = 5.0 / (H * W)          # at most a handful of threshold-flip pixels
    assert_close_tensor(st["color"].cpu()[sl], ref.color[sl], "color", 1e-4, budget)
    assert_close_tensor(st["depth"].cpu()[sl], ref.depth[sl], "depth", 1e-4, budget)
    assert_close_tensor(st["alpha"].cpu()[sl], ref.alpha[sl], "alpha", 1e-4, budget)
    assert_close_tensor(st["final_T"].cpu()[y0:y1], ref.img.final_T[y0:y1], "final_T", 1e-4, budget)
    mism = (st["n_contrib"].cpu()[y0:y1] != ref.img.n_contrib[y0:y1]).float().mean()
    assert float(mism) <= 2e-3, f"n_contrib mismatch fraction {float(mism):.2e}"
    if band is not None:      # pixels outside the band are left untouched (zero-initialised)
        assert float(st["color"].cpu()[:, :y0].abs().max()) == 0.0


def _oracle_grads(sc, cam, deg, bg, mod, grgb, touch_kw, gdepth=None, galpha=None, use_colors=False, use_cov=False):
    S = oracle_settings(cam, deg, bg, mod)
    ins = {k: v.clone().requires_grad_(True) for k, v in
           dict(means3D=sc.means3D, scales=sc.scales, rotations=sc.rotations, opacities=sc.opacities, shs=sc.shs).items()}
    kw = dict(shs=ins["shs"], scales=ins["scales"], rotations=ins["rotations"])
    if use_colors:
        pre0 = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, None, None, S)
        ins["colors"] = pre0.rgb.detach().clone().requires_grad_(True)
        kw["shs"] = None
        kw["colors_precomp"] = ins["colors"]
    if use_cov:
        pre0 = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, None, None, S)
        ins["cov3D"] = pre0.cov3D.detach().clone().requires_grad_(True)
        kw["scales"] = kw["rotations"] = None
        kw["cov3D_precomp"] = ins["cov3D"]
    out = O.rasterize(ins["means3D"], ins["opacities"], S, **kw, **touch_kw)
    loss = (out.color * grgb).sum() + out.touch_loss
    if gdepth is not None:
        loss = loss + (out.depth[0] * gdepth).sum()
    if galpha is not None:
        loss = loss + (out.alpha[0] * galpha).sum()
    loss.backward()
    return out, {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in ins.items()}


def _cuda_grads(sc, cam, deg, bg, mod, grgb, touch_kw, gdepth=None, galpha=None, colors=None, cov=None):
    rs = cuda_settings(cam, deg, DEV, bg, mod)
    ins = {k: v.to(DEV).clone().requires_grad_(True) for k, v in
           dict(means3D=sc.means3D, scales=sc.scales, rotations=sc.rotations, opacities=sc.opacities, shs=sc.shs).items()}
    kw = dict(shs=ins["shs"], scales=ins["scales"], rotations=ins["rotations"])
    if colors is not None:
        ins["colors"] = colors.to(DEV).clone().requires_grad_(True)
        kw["shs"] = None
        kw["colors_precomp"] = ins["colors"]
    if cov is not None:
        ins["cov3D"] = cov.to(DEV).clone().requires_grad_(True)
        kw["scales"] = kw["rotations"] = None
        kw["cov3D_precomp"] = ins["cov3D"]
    means2D = torch.zeros(sc.means3D.shape[0], 3, device=DEV, requires_grad=True)
    tk = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in touch_kw.items()}
    color, radii, depth, alpha, resid = T.GaussianRasterizer(rs)(ins["means3D"], means2D, ins["opacities"], **kw, **tk)
    loss = (color * grgb.to(DEV)).sum()
    if gdepth is not None:
        loss = loss + (depth[0] * gdepth.to(DEV)).sum()
    if galpha is not None:
        loss = loss + (alpha[0] * galpha.to(DEV)).sum()
    loss.backward()
    grads = {k: (v.grad.cpu() if v.grad is not None else torch.zeros_like(v).cpu()) for k, v in ins.items()}
    grads["means2D"] = means2D.grad.cpu()
    return (color, radii, depth, alpha, resid), grads


def _touch_inputs(sc, cam, deg, seed):
    S = oracle_settings(cam, deg)
    base = O.rasterize(sc.means3D, sc.opacities, S, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
    return synth.make_touch_maps(base.depth[0] + 0.03, seed=seed, n_patches=4, patch_radius=10)


GRAD_MODES = [
    dict(id="rgb_only", touch={}),
    dict(id="l1_norm", touch=dict(depth_loss="l1", depth_loss_mult=0.2, depth_normalize=True), need_touch=True),
    dict(id="l2_raw", touch=dict(depth_loss="l2", depth_loss_mult=0.5, depth_normalize=False), need_touch=True),
    dict(id="l1_fixed_norm_ext", touch=dict(depth_loss="l1", depth_loss_mult=0.005, depth_normalize=True,
                                              depth_loss_norm=1000.0), need_touch=True, ext=True),
    dict(id="ext_only_raw", touch=dict(depth_normalize=False), ext=True),
]


@pytest.mark.parametrize("mode", GRAD_MODES, ids=[m["id"] for m in GRAD_MODES])
@pytest.mark.parametrize("name", ["c1", "ragged", "close"])
def test_backward_parity(name, mode):
    c, sc, cam = _case(name)
    mod, bg = c.get("mod", 1.0), (0.2, 0.3, 0.1)
    H, W = c["H"], c["W"]
    g = torch.Generator().manual_seed(7)
    grgb = (torch.rand(3, H, W, generator=g) - 0.3) / (3 * H * W)
    touch = dict(mode["touch"])
    if mode.get("need_touch"):
        tgt, wgt = _touch_inputs(sc, cam, c["deg"], c["seed"])
        touch.update(touch_depth=tgt, touch_weight=wgt)
    gd = ga = None
    if mode.get("ext"):
        gd = torch.randn(H, W, generator=g) * 1e-4
        ga = torch.randn(H, W, generator=g) * 1e-4
    ref_out, ref = _oracle_grads(sc, cam, c["deg"], bg, mod, grgb, touch, gd, ga)
    out, got = _cuda_grads(sc, cam, c["deg"], bg, mod, grgb, touch, gd, ga)
    budget = 5.0 / (H * W)
    assert_close_tensor(out[0].cpu(), ref_out.color, "color", 1e-4, budget)
    assert_close_tensor(out[2].cpu(), ref_out.depth, "depth", 1e-4, budget)
    # residual = depth - target is a difference of nearly equal numbers: its error scales with |depth|
    dscale = float(ref_out.depth.abs().max())
    assert float((out[4].cpu() - ref_out.residual).abs().max()) <= 1e-4 * dscale or \
        float(((out[4].cpu() - ref_out.residual).abs() > 1e-4 * dscale).float().mean()) <= budget
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        assert_close_tensor(got[k], ref[k], "grad_" + k, 1e-4)
    assert torch.isfinite(got["means2D"]).all()


def test_touch_loss_honours_the_upstream_gradient():
    """the fused touch gradient must scale with whatever autograd carries.  (a) return_touch_loss=True:
    the sixth output is the differentiable loss value; (0.37 * (photo + touch)).backward() matches the oracle's
    0.37-scaled objective, and leaving the scalar out of the objective switches the fused gradient off.
    (b) injected mode with an explicit loss_grad_scale (GradScaler / accumulation)."""
    c, sc, cam = _case("c1")
    H, W = c["H"], c["W"]
    g = torch.Generator().manual_seed(17)
    grgb = torch.rand(3, H, W, generator=g) / (3 * H * W)
    tgt, wgt = _touch_inputs(sc, cam, c["deg"], 2)
    touch = dict(touch_depth=tgt, touch_weight=wgt, depth_loss="l1", depth_loss_mult=0.2)
    S = oracle_settings(cam, c["deg"])
    names = ("means3D", "scales", "rotations", "opacities", "shs")

    def oracle(scale_all, with_touch=True):
        ins = {k: getattr(sc, k).clone().requires_grad_(True) for k in names}
        out = O.rasterize(ins["means3D"], ins["opacities"], S, shs=ins["shs"], scales=ins["scales"],
                          rotations=ins["rotations"], **touch)
        loss = (out.color * grgb).sum() + (out.touch_loss if with_touch else 0.0)
        (scale_all * loss).backward()
        return out, {k: v.grad for k, v in ins.items()}

    def cuda(scale_all, mode):
        rs = cuda_settings(cam, c["deg"], DEV)
        ins = {k: getattr(sc, k).to(DEV).clone().requires_grad_(True) for k in names}
        tk = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in touch.items()}
        ras = T.GaussianRasterizer(rs)
        kw = dict(shs=ins["shs"], scales=ins["scales"], rotations=ins["rotations"])
        if mode == "returned":
            color, _, _, _, _, tl = ras(ins["means3D"], None, ins["opacities"], **kw, **tk, return_touch_loss=True)
            (scale_all * ((color * grgb.to(DEV)).sum() + tl)).backward()
        elif mode == "left_out":
            color, _, _, _, _, tl = ras(ins["means3D"], None, ins["opacities"], **kw, **tk, return_touch_loss=True)
            (scale_all * (color * grgb.to(DEV)).sum()).backward()
        else:
            out = ras(ins["means3D"], None, ins["opacities"], **kw, **tk,
                      loss_grad_scale=torch.tensor(scale_all, device=DEV))
            assert len(out) == 5
            tl = None
            (scale_all * (out[0] * grgb.to(DEV)).sum()).backward()
        return tl, {k: v.grad.cpu() for k, v in ins.items()}

    ref_out, ref = oracle(0.37)
    tl, got = cuda(0.37, "returned")
    assert abs(float(tl) - float(ref_out.touch_loss)) <= 1e-4 * abs(float(ref_out.touch_loss)) + 1e-9
    for k in names:
        assert_close_tensor(got[k], ref[k], "grad_" + k + " (returned touch loss x0.37)", 1e-4)
    _, got = cuda(0.37, "injected")
    for k in names:
        assert_close_tensor(got[k], ref[k], "grad_" + k + " (loss_grad_scale 0.37)", 1e-4)
    _, ref0 = oracle(0.37, with_touch=False)
    _, got = cuda(0.37, "left_out")
    for k in names:
        assert_close_tensor(got[k], ref0[k], "grad_" + k + " (touch loss left out)", 1e-4)
    assert rel_inf(ref0["means3D"], ref["means3D"]) > 1e-2, "the touch term must matter in this scene"


def test_backward_precomputed_colors_and_cov():
    c, sc, cam = _case("c1")
    H, W = c["H"], c["W"]
    g = torch.Generator().manual_seed(8)
    grgb = torch.rand(3, H, W, generator=g) / (3 * H * W)
    tgt, wgt = _touch_inputs(sc, cam, c["deg"], 1)
    touch = dict(touch_depth=tgt, touch_weight=wgt, depth_loss="l1", depth_loss_mult=0.2)
    ref_out, ref = _oracle_grads(sc, cam, c["deg"], (0, 0, 0), 1.0, grgb, touch, use_colors=True, use_cov=True)
    out, got = _cuda_grads(sc, cam, c["deg"], (0, 0, 0), 1.0, grgb, touch,
                           colors=ref_out.pre.rgb.detach(), cov=ref_out.pre.cov3D.detach())
    assert_close_tensor(out[0].cpu(), ref_out.color, "color", 1e-4, 5.0 / (H * W))
    for k in ("means3D", "opacities", "colors", "cov3D"):
        assert_close_tensor(got[k], ref[k], "grad_" + k, 1e-4)


def test_screen_space_gradients_via_c_abi():
    """tgs_backward_render's [N,10] buffer (the all-reduced quantity) against oracle partials."""
    c, sc, cam = _case("c1")
    H, W, N = c["H"], c["W"], c["N"]
    S = oracle_settings(cam, c["deg"], (0.1, 0.2, 0.3))
    pre = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, None, None, S)
    leaves = {k: getattr(pre, k).detach().clone().requires_grad_(True) for k in ("xy", "conic", "opacity", "rgb", "depth")}
    lpre = pre._replace(**leaves)
    img = O.render_tiles(lpre, O.bin_and_sort(lpre, S), S)
    g = torch.Generator().manual_seed(9)
    grgb = torch.rand(3, H, W, generator=g)
    gdep = torch.rand(H, W, generator=g) * 0.1
    galp = torch.rand(H, W, generator=g) * 0.1
    ((img.color * grgb).sum() + (img.depth * gdep).sum() + (img.alpha * galp).sum()).backward()
    ref = torch.cat([leaves["xy"].grad, leaves["conic"].grad, leaves["opacity"].grad[:, None],
                     leaves["rgb"].grad, leaves["depth"].grad[:, None]], 1)

    lib, L = T._lib.load(), T._lib
    rs = cuda_settings(cam, c["deg"], DEV, (0.1, 0.2, 0.3))
    m, s, r, o, sh = _to(DEV, sc.means3D, sc.scales, sc.rotations, sc.opacities.reshape(-1), sc.shs)
    keep = []
    opt = T.TouchOptions(depth_normalize=False)
    from importlib import import_module
    R = import_module("touch-gs_b200.rasterizer")
    st, _ = R._make_settings(rs, opt, 1, keep)
    gs = R._make_gaussians(m, o, sh, None, s, r, None)
    color = torch.empty(3, H, W, device=DEV); depth = torch.empty(H, W, device=DEV); alpha = torch.empty(H, W, device=DEV)
    radii = torch.zeros(N, dtype=torch.int32, device=DEV)
    scratch = R._Scratch(DEV)
    saved = L.TgsSaved()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    L.check(lib.tgs_forward(C.byref(st), C.byref(gs), scratch.cb, None, p(color), p(depth), p(alpha), p(radii),
                            None, None, C.byref(saved), stream), "fwd")
    sg = torch.full((N, 10), 123.0, device=DEV)          # must be zeroed by the library
    gr, gd, ga = grgb.to(DEV), gdep.to(DEV), galp.to(DEV)
    L.check(lib.tgs_backward_render(C.byref(st), C.byref(gs), C.byref(saved), p(gr), p(gd), p(ga), None, None,
                                    p(sg), stream), "bwd_render")
    torch.cuda.synchronize()
    assert_close_tensor(sg.cpu(), ref, "screen_grads", 1e-4)
    for col in range(10):
        assert_close_tensor(sg.cpu()[:, col], ref[:, col], f"screen_grads[:, {col}]", 2e-4)


def test_golden_fixture():
    """CUDA path against the committed fixture tests/golden/oracle_c1_small.npz."""
    from golden.make_golden import golden_inputs
    z = np.load(os.path.join(ROOT, "tests", "golden", "oracle_c1_small.npz"))
    sc, cam, grgb, target, weight = golden_inputs()
    bg = (0.2, 0.1, 0.3)
    rs = cuda_settings(cam, sc.sh_degree, DEV, bg)
    m, s, r, o, sh = _to(DEV, sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)
    st = T.inspect_state.forward_state(m, o, rs, shs=sh, scales=s, rotations=r)
    assert np.array_equal(st["radii"].cpu().numpy(), z["radii"])
    assert np.array_equal(st["keys"].cpu().numpy(), z["keys"])
    assert np.array_equal(st["vals"].cpu().numpy(), z["vals"])
    assert np.array_equal(st["ranges"].cpu().numpy(), z["ranges"])
    assert float((st["n_contrib"].cpu() != torch.from_numpy(z["n_contrib"])).float().mean()) <= 2e-3
    touch = dict(touch_depth=target, touch_weight=weight, depth_loss="l1", depth_loss_mult=0.2, depth_normalize=True)
    out, got = _cuda_grads(sc, cam, sc.sh_degree, bg, 1.0, grgb, touch)
    budget = 5.0 / (48 * 64)
    assert_close_tensor(out[0].cpu(), torch.from_numpy(z["color"]), "color", 1e-4, budget)
    assert_close_tensor(out[2].cpu(), torch.from_numpy(z["depth"]), "depth", 1e-4, budget)
    assert float((out[4].cpu() - torch.from_numpy(z["residual"])).abs().max()) <= 1e-4 * float(np.abs(z["depth"]).max())
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        assert_close_tensor(got[k], torch.from_numpy(z["grad_" + k]), "grad_" + k, 1e-4)


def test_deferred_count_is_identical_and_overflow_raises_in_backward():
    """defer_count: the forward never waits for num_rendered; with a sufficient hint nothing changes, with a hint that
    is too small the backward refuses to run on the truncated forward."""
    c, sc, cam = _case("ragged")
    rs = cuda_settings(cam, c["deg"], DEV, (0.1, 0.2, 0.3), c.get("mod", 1.0))
    m, s, r, o, sh = _to(DEV, sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)
    H, W = c["H"], c["W"]
    g = torch.Generator().manual_seed(2)
    grgb = (torch.rand(3, H, W, generator=g) / (3 * H * W)).to(DEV)

    def run(**kw):
        mm = m.clone().requires_grad_(True)
        ras = T.GaussianRasterizer(rs)
        out = ras(mm, None, o, shs=sh, scales=s, rotations=r, **kw)
        (out[0] * grgb).sum().backward()
        return out, mm.grad, ras
    ref, gref, ras0 = run()
    I = ras0.last_num_rendered
    for binding in ("ext", "ctypes"):
        T._lib.use_binding(binding)
        try:
            out, grad, ras = run(rendered_hint=int(I * 1.2), defer_count=True)
            for a, b in zip(out[:5], ref[:5]):
                assert torch.equal(a, b)
            assert rel_inf(grad, gref) < 1e-5
            assert ras.last_num_rendered == I                      # the ticket is redeemed on access
            with pytest.raises(Exception, match="overflowed"):
                run(rendered_hint=max(1, int(I * 0.3)), defer_count=True)
            out2, grad2, _ = run(rendered_hint=int(I * 0.3))         # without defer_count: exact re-run, same result
            assert torch.equal(out2[0], ref[0]) and rel_inf(grad2, gref) < 1e-5
        finally:
            T._lib.use_binding("ext")


# ----------------------------------------------------------------------------- edge cases
def test_empty_input():
    cam = synth.look_at_camera(64, 48, (0, 0, -3.0))
    rs = cuda_settings(cam, 0, DEV, (0.5, 0.25, 0.125))
    m = torch.zeros(0, 3, device=DEV, requires_grad=True)
    out = T.GaussianRasterizer(rs)(m, None, torch.zeros(0, 1, device=DEV), shs=torch.zeros(0, 1, 3, device=DEV),
                                   scales=torch.zeros(0, 3, device=DEV), rotations=torch.zeros(0, 4, device=DEV))
    color, radii, depth, alpha, resid = out
    assert radii.numel() == 0
    assert torch.allclose(color[0], torch.full((48, 64), 0.5, device=DEV)) and float(alpha.abs().max()) == 0.0
    color.sum().backward()
    assert m.grad.shape == (0, 3)


def test_all_culled_and_behind_camera():
    cam = synth.look_at_camera(64, 64, (0, 0, -3.0))
    rs = cuda_settings(cam, 0, DEV)
    sc = synth.make_scene(200, 0, 0.02, 0.1, seed=3)
    m = (sc.means3D + torch.tensor([0.0, 0.0, -10.0])).to(DEV).requires_grad_(True)     # all behind the camera
    s, r, o, sh = _to(DEV, sc.scales, sc.rotations, sc.opacities, sc.shs)
    color, radii, depth, alpha, resid = T.GaussianRasterizer(rs)(m, None, o, shs=sh, scales=s, rotations=r)
    assert int(radii.max()) == 0 and float(color.abs().max()) == 0.0
    (color.sum() + depth.sum()).backward()
    assert float(m.grad.abs().max()) == 0.0
    vis = T.GaussianRasterizer(rs).markVisible(m.detach())
    assert int(vis.sum()) == 0


def test_mark_visible_matches_oracle():
    cam = synth.look_at_camera(64, 64, (0.2, 0.1, -0.5))
    sc = synth.make_scene(5000, 0, 0.02, 0.1, seed=6)
    rs = cuda_settings(cam, 0, DEV)
    got = T.GaussianRasterizer(rs).markVisible(sc.means3D.to(DEV)).cpu()
    ref = O.mark_visible(sc.means3D, cam.viewmatrix)
    assert torch.equal(got, ref) and 0 < int(ref.sum()) < 5000


def test_single_huge_gaussian_covers_everything():
    cam = synth.look_at_camera(100, 60, (0, 0, -3.0))
    S = oracle_settings(cam, 0, (0.0, 0.0, 0.0))
    m = torch.tensor([[0.0, 0.0, 0.0]]); s = torch.tensor([[5.0, 5.0, 5.0]]); q = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    o = torch.tensor([[0.9]]); sh = torch.tensor([[[1.0, 0.5, -0.2]]])
    ref = O.rasterize(m, o, S, shs=sh, scales=s, rotations=q)
    rs = cuda_settings(cam, 0, DEV)
    color, radii, depth, alpha, _ = T.GaussianRasterizer(rs)(*_to(DEV, m), None, *_to(DEV, o), shs=sh.to(DEV),
                                                             scales=s.to(DEV), rotations=q.to(DEV))
    assert int(radii[0]) == int(ref.radii[0]) > 100
    assert_close_tensor(color.cpu(), ref.color, "color", 1e-4)
    assert_close_tensor(depth.cpu(), ref.depth, "depth", 1e-4)
    assert float(alpha.min()) > 0.5


def test_host_buffer_entry_point_matches_operator():
    """tgs_train_step_host (HOST pointers in, HOST pointers out) == operator path with the same L1 loss."""
    c, sc, cam = _case("c1")
    H, W, N = c["H"], c["W"], c["N"]
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(3, H, W, generator=g)
    tgt, wgt = _touch_inputs(sc, cam, c["deg"], 2)
    bg = torch.tensor([0.1, 0.1, 0.1])
    lib, L = T._lib.load(), T._lib
    p = lambda t: C.c_void_p(t.data_ptr())
    vm, pmx, cp = cam.viewmatrix.contiguous(), cam.projmatrix.contiguous(), cam.campos.contiguous()
    s = L.TgsSettings(image_width=W, image_height=H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, scale_modifier=1.0,
                      sh_degree=0, sh_coeffs=1, prefiltered=0, debug=0, tile_row_begin=0, tile_row_end=0,
                      depth_normalize=1, viewmatrix=vm.data_ptr(), projmatrix=pmx.data_ptr(), campos=cp.data_ptr(),
                      bg=bg.data_ptr())
    op = sc.opacities.reshape(-1).contiguous()
    gs = L.TgsGaussians(N=N, means3D=sc.means3D.data_ptr(), opacities=op.data_ptr(), shs=sc.shs.data_ptr(),
                        colors_precomp=None, scales=sc.scales.data_ptr(), rotations=sc.rotations.data_ptr(),
                        cov3D_precomp=None)
    d = {k: torch.zeros(sh_) for k, sh_ in dict(m2=(N, 3), m3=(N, 3), o=(N,), sh=(N, 1, 3), s=(N, 3), r=(N, 4)).items()}
    gr = L.TgsGrads(dmeans2D=d["m2"].data_ptr(), dmeans3D=d["m3"].data_ptr(), dopacity=d["o"].data_ptr(),
                    dshs=d["sh"].data_ptr(), dcolors=None, dscales=d["s"].data_ptr(), drotations=d["r"].data_ptr(),
                    dcov3D=None)
    color_h = torch.zeros(3, H, W); depth_h = torch.zeros(H, W); radii_h = torch.zeros(N, dtype=torch.int32)
    loss_h = torch.zeros(1); nr = C.c_int64(0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.tgs_train_step_host(C.byref(s), C.byref(gs), p(gt), p(tgt), p(wgt), L.LOSS_L1, 0.2, C.byref(gr),
                                    p(color_h), p(depth_h), p(radii_h), p(loss_h), C.byref(nr), stream), "host step")
    # operator path
    rs = cuda_settings(cam, 0, DEV, (0.1, 0.1, 0.1))
    ins = {k: v.to(DEV).clone().requires_grad_(True) for k, v in
           dict(m=sc.means3D, s=sc.scales, r=sc.rotations, o=sc.opacities, sh=sc.shs).items()}
    color, radii, depth, alpha, resid = T.GaussianRasterizer(rs)(
        ins["m"], None, ins["o"], shs=ins["sh"], scales=ins["s"], rotations=ins["r"],
        touch_depth=tgt.to(DEV), touch_weight=wgt.to(DEV), depth_loss="l1", depth_loss_mult=0.2)
    loss = (color - gt.to(DEV)).abs().mean()
    loss.backward()
    assert nr.value > 0 and torch.equal(radii.cpu(), radii_h)
    assert torch.equal(color.detach().cpu(), color_h) and torch.equal(depth.detach().cpu()[0], depth_h)
    assert abs(float(loss) - float(loss_h)) < 1e-5
    for a, b in ((d["m3"], ins["m"].grad), (d["s"], ins["s"].grad), (d["r"], ins["r"].grad),
                 (d["o"], ins["o"].grad.reshape(-1)), (d["sh"], ins["sh"].grad)):
        assert_close_tensor(a, b.cpu(), "host-step grad", 1e-4)


# ------------------------------------------------------------- full-size property tests (c3)
@pytest.fixture(scope="module")
def c3_state():
    cfg = synth.CONFIGS["c3"]
    sc = synth.make_scene(cfg["N"], cfg["sh_degree"], cfg["smin"], cfg["smax"], seed=0)
    cam = synth.orbit_cameras(cfg["W"], cfg["H"], 8, 3.0, 0)[0]
    rs = cuda_settings(cam, cfg["sh_degree"], DEV)
    t = _to(DEV, sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)
    return cfg, cam, rs, t


def test_fullsize_integer_invariants(c3_state):
    cfg, cam, rs, (m, s, r, o, sh) = c3_state
    st = T.inspect_state.forward_state(m, o, rs, shs=sh, scales=s, rotations=r)
    I = st["num_rendered"]
    assert I == int(st["tiles_touched"].long().sum()) == st["num_rendered_device"] > 1_000_000
    keys = st["keys"]
    assert bool((keys[1:] >= keys[:-1]).all()), "sorted keys not monotone"
    # the list as a multiset == every (tile of its rectangle, Gaussian) pair exactly once
    fin = (st["tile_ids"] << 32) | st["vals"].long()
    vis_ids = torch.nonzero(st["tiles_touched"] > 0).flatten()
    x0, y0 = st["rect_min"][vis_ids, 0].long(), st["rect_min"][vis_ids, 1].long()
    w = st["rect_max"][vis_ids, 0].long() - x0
    cnt = st["tiles_touched"][vis_ids].long()
    gid = torch.repeat_interleave(vis_ids, cnt)
    loc = torch.arange(I, device=gid.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
    ww, xx0, yy0 = (torch.repeat_interleave(t, cnt) for t in (w, x0, y0))
    Tx = (cfg["W"] + 15) // 16
    em = (((yy0 + loc // ww) * Tx + xx0 + loc % ww) << 32) | gid
    assert torch.equal(torch.sort(em).values, torch.sort(fin).values), "the list is not the multiset of rectangle instances"
    same = keys[1:] == keys[:-1]
    assert bool((st["vals"][1:][same] >= st["vals"][:-1][same]).all()), "sort not stable"
    rg = st["ranges"].long()
    nz = rg[:, 1] > rg[:, 0]
    assert int((rg[nz, 1] - rg[nz, 0]).sum()) == I, "ranges do not partition the list"
    tiles = keys >> 32
    starts = rg[nz, 0]
    assert torch.equal(tiles[starts], torch.nonzero(nz).flatten()), "range start does not match its tile"
    assert bool(((st["radii"] > 0) == (st["tiles_touched"] > 0)).all())
    fT, A = st["final_T"], st["alpha"][0]
    assert float(fT.min()) >= 0.0 and float(fT.max()) <= 1.0
    assert float((A + fT - 1.0).abs().max()) < 1e-6
    assert int(st["n_contrib"].max()) <= int((rg[:, 1] - rg[:, 0]).max())
    assert torch.isfinite(st["color"]).all() and torch.isfinite(st["depth"]).all()
    # expected depth lies inside the scene's depth range wherever anything was hit
    hit = A > 0.5
    d = st["depth"][0][hit]
    assert float(d.min()) > 0.2 and float(d.max()) < 6.0


def test_fullsize_determinism_linearity_and_bands(c3_state):
    cfg, cam, rs, (m, s, r, o, sh) = c3_state
    H, W = cfg["H"], cfg["W"]
    ras = T.GaussianRasterizer(rs)
    g = torch.Generator().manual_seed(0)
    grgb = (torch.rand(3, H, W, generator=g) / (3 * H * W)).to(DEV)

    def run(scale, tile_rows=None):
        mm = m.clone().requires_grad_(True)
        oo = o.clone().requires_grad_(True)
        color, radii, depth, alpha, _ = ras(mm, None, oo, shs=sh, scales=s, rotations=r, tile_rows=tile_rows)
        (color * grgb * scale).sum().backward()
        return color.detach(), depth.detach(), mm.grad, oo.grad, radii

    c1, d1, gm1, go1, rad1 = run(1.0)
    c2, d2, gm2, go2, _ = run(1.0)
    assert torch.equal(c1, c2) and torch.equal(d1, d2), "forward is not deterministic"
    assert rel_inf(gm2, gm1) < 1e-4                      # atomics reorder sums: tolerance, not bits
    c3, _, gm3, go3, _ = run(2.0)
    assert rel_inf(gm3, 2.0 * gm1) < 1e-4 and rel_inf(go3, 2.0 * go1) < 1e-4, "backward is not linear in dL/dcolor"
    # tile-row bands: band renders are bit-identical slices; partial grads sum to the full
    bands = T.sharding.even_bands(H, 2)
    acc_m, acc_o = torch.zeros_like(gm1), torch.zeros_like(go1)
    for b in bands:
        cb, db, gmb, gob, radb = run(1.0, b)
        y0, y1 = T.sharding.band_pixel_rows(b, H)
        assert torch.equal(cb[:, y0:y1], c1[:, y0:y1]) and torch.equal(db[:, y0:y1], d1[:, y0:y1])
        assert torch.equal(radb, rad1)
        acc_m += gmb
        acc_o += gob
    assert rel_inf(acc_m, gm1) < 1e-4 and rel_inf(acc_o, go1) < 1e-4


@pytest.mark.parametrize("name", ["c1", "ragged", "band"])        # c1: T = 64 = 2^6 exercises the pad-key bit
@pytest.mark.parametrize("hint_scale", [0.3, 1.0, 1.7])           # overflow (exact re-run) / exact / slack
def test_speculative_sizing_is_bit_identical(name, hint_scale):
    """rendered_hint only changes WHEN the host learns num_rendered, never the result."""
    c, sc, cam = _case(name)
    band = c.get("band")
    rs = cuda_settings(cam, c["deg"], DEV, (0.1, 0.2, 0.3), c.get("mod", 1.0))
    m, s, r, o, sh = _to(DEV, sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)
    ref = T.inspect_state.forward_state(m, o, rs, shs=sh, scales=s, rotations=r, opt=T.TouchOptions(tile_rows=band))
    I = ref["num_rendered"]
    hint = max(1, int(I * hint_scale))
    got = T.inspect_state.forward_state(m, o, rs, shs=sh, scales=s, rotations=r,
                                        opt=T.TouchOptions(tile_rows=band, rendered_hint=hint))
    assert got["num_rendered"] == I
    for k in ("color", "depth", "alpha", "radii", "ranges"):
        assert torch.equal(got[k], ref[k]), k
    y0, y1 = (0, c["H"]) if band is None else T.sharding.band_pixel_rows(band, c["H"])
    for k in ("final_T", "n_contrib"):            # per-pixel saved state exists only inside the band
        assert torch.equal(got[k][y0:y1], ref[k][y0:y1]), k
    assert torch.equal(got["keys"][:I], ref["keys"][:I]) and torch.equal(got["vals"][:I], ref["vals"][:I])
    assert torch.equal(got["records"][:I], ref["records"][:I])
    # and the backward through the operator agrees too
    H, W = c["H"], c["W"]
    g = torch.Generator().manual_seed(2)
    grgb = (torch.rand(3, H, W, generator=g) / (3 * H * W)).to(DEV)

    def grads(h):
        mm = m.clone().requires_grad_(True)
        ras = T.GaussianRasterizer(rs)
        color = ras(mm, None, o, shs=sh, scales=s, rotations=r, tile_rows=band, rendered_hint=h)[0]
        (color * grgb).sum().backward()
        assert ras.last_num_rendered == I
        return mm.grad
    assert rel_inf(grads(hint), grads(0)) < 1e-5


def test_config_c2_full_parity():
    """BASELINE config c2 at FULL size (100k Gaussians, 800x800, SH degree 3) with the fused touch depth-L1
    loss: integer state bit-exact, images and all five gradient tensors within 1e-4 of the oracle."""
    cfg = synth.CONFIGS["c2"]
    sc = synth.make_scene(cfg["N"], cfg["sh_degree"], cfg["smin"], cfg["smax"], seed=0)
    cam = synth.orbit_cameras(cfg["W"], cfg["H"], 8, 3.0, 0)[0]
    H, W = cfg["H"], cfg["W"]
    g = torch.Generator().manual_seed(12)
    grgb = (torch.rand(3, H, W, generator=g) - 0.5) / (3 * H * W)
    # touch target from a depth render of the perturbed scene by the CUDA path itself (as bench.py does)
    rs = cuda_settings(cam, cfg["sh_degree"], DEV)
    pert = synth.perturbed(sc, 0.01, 0)
    with torch.no_grad():
        d = T.GaussianRasterizer(rs)(*_to(DEV, pert.means3D), None, *_to(DEV, pert.opacities), shs=pert.shs.to(DEV),
                                     scales=pert.scales.to(DEV), rotations=pert.rotations.to(DEV))[2]
    tgt, wgt = synth.make_touch_maps(d[0].cpu(), seed=0)
    touch = dict(touch_depth=tgt, touch_weight=wgt, depth_loss="l1", depth_loss_mult=0.2, depth_normalize=True)
    ref_out, ref = _oracle_grads(sc, cam, cfg["sh_degree"], (0, 0, 0), 1.0, grgb, touch)
    m, s, r, o, sh = _to(DEV, sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs)
    st = T.inspect_state.forward_state(m, o, rs, shs=sh, scales=s, rotations=r)
    assert torch.equal(st["radii"].cpu(), ref_out.radii)
    assert torch.equal(st["keys"].cpu(), ref_out.bins.keys) and torch.equal(st["vals"].cpu(), ref_out.bins.vals)
    assert torch.equal(st["ranges"].cpu(), ref_out.bins.ranges)
    assert float((st["n_contrib"].cpu() != ref_out.img.n_contrib).float().mean()) <= 2e-3
    out, got = _cuda_grads(sc, cam, cfg["sh_degree"], (0, 0, 0), 1.0, grgb, touch)
    budget = 20.0 / (H * W)
    assert_close_tensor(out[0].cpu(), ref_out.color, "color", 1e-4, budget)
    assert_close_tensor(out[2].cpu(), ref_out.depth, "depth", 1e-4, budget)
    assert_close_tensor(out[3].cpu(), ref_out.alpha, "alpha", 1e-4, budget)
    for k in ("means3D", "scales", "rotations", "opacities", "shs"):
        assert_close_tensor(got[k], ref[k], "grad_" + k, 1e-4)
