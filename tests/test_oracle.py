"""CPU tests of the oracle itself (the oracle is parity-UNPINNED by the reference: it has no tests,
golden vectors or fixtures for this path -- SURVEY.md §4, §8c -- so these are self-consistency checks
plus a committed regression fixture)."""
import os

import numpy as np
import pytest
import torch

from helpers import O, synth, oracle_settings, rel_inf, ROOT


def _run(scene, cam, dt, naive=False, touch=None, band=None, bg=(0.1, 0.2, 0.3), **kw):
    S = oracle_settings(cam, scene.sh_degree, bg)
    ins = [t.to(dt).clone().requires_grad_(True) for t in
           (scene.means3D, scene.scales, scene.rotations, scene.opacities, scene.shs)]
    m, s, r, o, sh = ins
    out = O.rasterize(m, o, S, shs=sh, scales=s, rotations=r, naive=naive, band=band, **(touch or {}), **kw)
    return out, ins


def _loss(out, H, W, dt, seed=3):
    g = torch.Generator().manual_seed(seed)
    grgb = torch.rand(3, H, W, generator=g).to(dt)
    return (out.color * grgb).sum() + (out.depth * 0.3).sum() + (out.alpha * 0.2).sum() + out.touch_loss


def test_naive_equals_vectorised():
    sc = synth.make_scene(60, 2, 0.05, 0.3, seed=1)
    cam = synth.look_at_camera(40, 24, (0.5, 0.3, -3.0))
    a, ia = _run(sc, cam, torch.float64, naive=False)
    b, ib = _run(sc, cam, torch.float64, naive=True)
    assert torch.equal(a.img.n_contrib, b.img.n_contrib)
    assert rel_inf(a.color, b.color) < 1e-12 and rel_inf(a.depth, b.depth) < 1e-12
    _loss(a, 24, 40, torch.float64).backward()
    _loss(b, 24, 40, torch.float64).backward()
    for x, y in zip(ia, ib):
        assert rel_inf(x.grad, y.grad) < 1e-10


def test_fp32_vs_fp64():
    sc = synth.make_scene(300, 3, 0.03, 0.2, seed=2)
    cam = synth.look_at_camera(64, 48, (0.4, -0.2, -3.0))
    tgt = torch.full((48, 64), 2.9)
    touch = dict(touch_depth=tgt, touch_weight=torch.ones(48, 64), depth_loss="l2", depth_loss_mult=0.2)
    a, ia = _run(sc, cam, torch.float32, touch=touch)
    b, ib = _run(sc, cam, torch.float64, touch=touch)
    assert rel_inf(a.color, b.color) < 1e-5
    _loss(a, 48, 64, torch.float32).backward()
    _loss(b, 48, 64, torch.float64).backward()
    for x, y in zip(ia, ib):
        assert rel_inf(x.grad, y.grad) < 2e-4


def test_autograd_vs_finite_differences():
    """fp64 central differences on a handful of coordinates (the render is piecewise smooth; the
    step is small enough that no alpha/T threshold is crossed for this seed)."""
    sc = synth.make_scene(12, 1, 0.1, 0.4, seed=5)
    cam = synth.look_at_camera(32, 32, (0.2, 0.1, -3.0))
    tgt = torch.full((32, 32), 3.0, dtype=torch.float64)
    touch = dict(touch_depth=tgt, touch_weight=None, depth_loss="l2", depth_loss_mult=0.5)
    out, ins = _run(sc, cam, torch.float64, touch=touch)
    _loss(out, 32, 32, torch.float64).backward()
    S = oracle_settings(cam, sc.sh_degree, (0.1, 0.2, 0.3))
    base = [t.detach().clone() for t in ins]

    def f(vals):
        m, s, r, o, sh = vals
        return float(_loss(O.rasterize(m, o, S, shs=sh, scales=s, rotations=r, **touch), 32, 32, torch.float64))
    h = 1e-6
    for ti, idx in [(0, (3, 0)), (0, (7, 2)), (1, (2, 1)), (2, (5, 3)), (3, (4, 0)), (4, (1, 2, 1))]:
        p = [t.clone() for t in base]
        q = [t.clone() for t in base]
        p[ti][idx] += h
        q[ti][idx] -= h
        fd = (f(p) - f(q)) / (2 * h)
        an = float(ins[ti].grad[idx])
        assert abs(fd - an) <= 1e-4 * max(1.0, abs(an)), (ti, idx, fd, an)


def test_binning_properties():
    sc = synth.make_scene(2000, 0, 0.02, 0.2, seed=0)
    cam = synth.look_at_camera(128, 128, (0.5, 0.3, -3.0))
    S = oracle_settings(cam, 0)
    pre = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, None, None, S)
    bins = O.bin_and_sort(pre, S)
    I = int(pre.tiles_touched.sum())
    assert bins.keys.numel() == I == int(bins.offsets[-1])
    assert bool((bins.keys[1:] >= bins.keys[:-1]).all())
    r = bins.ranges
    nz = r[:, 1] > r[:, 0]
    assert int((r[nz, 1] - r[nz, 0]).sum()) == I          # ranges partition the list
    tiles = (bins.keys >> 32)
    for t in torch.nonzero(nz).flatten()[:20].tolist():
        assert bool((tiles[r[t, 0]:r[t, 1]] == t).all())
    # stability: equal keys keep ascending Gaussian id
    same = bins.keys[1:] == bins.keys[:-1]
    assert bool((bins.vals[1:][same] >= bins.vals[:-1][same]).all())
    assert (pre.radii > 0).sum() > 0 and bool(((pre.radii > 0) == (pre.tiles_touched > 0)).all())


def test_band_partition_is_exact():
    """Rendering tile-row bands separately reproduces the full image exactly."""
    sc = synth.make_scene(800, 1, 0.03, 0.25, seed=4)
    cam = synth.look_at_camera(96, 80, (0.3, 0.2, -3.0))
    full, _ = _run(sc, cam, torch.float32)
    parts = [_run(sc, cam, torch.float32, band=b)[0] for b in ((0, 2), (2, 5))]
    assert torch.equal(full.color[:, :32], parts[0].color[:, :32])
    assert torch.equal(full.color[:, 32:80], parts[1].color[:, 32:80])
    assert torch.equal(full.radii, parts[0].radii) and torch.equal(full.radii, parts[1].radii)
    assert int(parts[0].pre.tiles_touched.sum() + parts[1].pre.tiles_touched.sum()) == int(full.pre.tiles_touched.sum())


def test_touch_loss_semantics():
    d = torch.tensor([[2.0, 0.0], [1.0, 3.0]])
    a = torch.tensor([[0.5, 0.0], [1.0, 0.5]])
    tgt = torch.tensor([[3.0, 1.0], [0.0, 5.0]])      # (1,0) invalid target, (0,1) no coverage
    w = torch.tensor([[2.0, 1.0], [1.0, 1.0]])
    sc = O.loss_scale_from_target(tgt, 0.2)
    assert abs(sc - 0.2 / 3) < 1e-12
    l1, r, dh = O.touch_loss(d, a, tgt, w, "l1", sc, True)
    assert torch.allclose(dh, torch.tensor([[4.0, 0.0], [1.0, 6.0]]))
    assert torch.allclose(r, torch.tensor([[1.0, 0.0], [0.0, 1.0]]))
    assert abs(float(l1) - sc * (2.0 * 1.0 + 1.0 * 1.0)) < 1e-7
    l2, _, _ = O.touch_loss(d, a, tgt, w, "l2", sc, False)
    assert abs(float(l2) - sc * (2.0 * 1.0 + 1.0 * 4.0)) < 1e-6


GOLDEN = os.path.join(ROOT, "tests", "golden", "oracle_c1_small.npz")


def test_golden_regression_fixture():
    """Committed fixture produced by tests/golden/make_golden.py from the oracle (there are no
    reference-side vectors to pin against): guards the oracle against silent drift."""
    z = np.load(GOLDEN)
    from golden.make_golden import golden_case
    out, grads = golden_case()
    assert np.array_equal(z["radii"], out.radii.numpy())
    assert np.array_equal(z["keys"], out.bins.keys.numpy())
    assert np.array_equal(z["vals"], out.bins.vals.numpy())
    assert np.array_equal(z["ranges"], out.bins.ranges.numpy())
    assert np.array_equal(z["n_contrib"], out.img.n_contrib.numpy())
    np.testing.assert_allclose(z["color"], out.color.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(z["depth"], out.depth.detach().numpy(), rtol=1e-5, atol=1e-6)
    for k, g in grads.items():
        np.testing.assert_allclose(z["grad_" + k], g.numpy(), rtol=2e-4, atol=1e-6 * float(np.abs(z["grad_" + k]).max()))


def test_fixture_scene_and_zbuffer_match_reference_function():
    """the fixture loader, and the z-buffer depth target against golden vectors produced by the
    reference's own ``project_points_with_colors`` (tests/golden/make_zbuffer_golden.py)."""
    import os
    import numpy as np
    from helpers import ROOT, T
    sc = T.synth.fixture_scene(0)
    assert sc.means3D.shape == (71283, 3) and sc.shs.shape == (71283, 1, 3)
    assert float(sc.scales.min()) > 0 and torch.allclose(sc.rotations.norm(dim=-1), torch.ones(71283), atol=1e-5)
    big = T.synth.fixture_scene(1, copies=3)
    assert big.means3D.shape[0] == 3 * 71283 and torch.equal(big.means3D[:71283], sc.means3D)
    z = np.load(os.path.join(ROOT, "tests", "golden", "zbuffer_reference.npz"))
    W, H = 160, 120
    cams = T.synth.fixture_cameras(W, H, 3)
    for i, cam in enumerate(cams):
        got = T.synth.zbuffer_depth(sc.means3D[::7], cam)
        ref = torch.from_numpy(z[f"depth_{i}"])
        assert got.shape == ref.shape and float((ref > 0).float().mean()) > 0.01
        assert torch.equal(got > 0, ref > 0), "z-buffer coverage differs from the reference function"
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6)
        # every camera looks at the common focus of the 100 sample poses
        f = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "fixture_scene.npz"))["focus"])
        v = cam.viewmatrix.t() @ torch.cat([f, torch.ones(1)])
        assert abs(float(v[0])) < 1e-4 and abs(float(v[1])) < 1e-4 and 0.3 < float(v[2]) < 0.8


def test_oracle_convention_switches_are_self_consistent():
    """SURVEY A.3 switches of the oracle (used by the gsplat-style surface tests): sampling at (x+0.5, y+0.5) equals
    integer sampling of splats moved by -0.5; the principal point shifts the pixel means by exactly that many pixels;
    near_z moves the cull plane; alpha_max only matters where o*G exceeds the smaller clamp."""
    from helpers import O, synth
    sc = synth.make_scene(400, 0, 0.03, 0.3, seed=21)
    cam = synth.look_at_camera(96, 80, (0.2, 0.1, -2.4))
    g = torch.Generator().manual_seed(0)
    colors = torch.rand(400, 3, generator=g)
    base = dict(image_height=80, image_width=96, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.zeros(3), scale_modifier=1.0,
                viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, sh_degree=0, campos=cam.campos)
    S0 = O.OracleSettings(**base)
    S1 = O.OracleSettings(**base, pixel_offset=0.5)
    pre = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, None, colors, None, S0)
    bins = O.bin_and_sort(pre, S0)
    a = O.render_tiles(pre, bins, S1)
    b = O.render_tiles(pre._replace(xy=pre.xy - 0.5), bins, S0)
    assert torch.allclose(a.color, b.color, atol=1e-6) and torch.equal(a.n_contrib, b.n_contrib)
    Sp = O.OracleSettings(**base, principal=(3.25, -2.5))
    prep = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, None, colors, None, Sp)
    vis = (pre.radii > 0) & (prep.radii > 0)
    assert torch.allclose(prep.xy[vis] - pre.xy[vis], torch.tensor([3.25, -2.5]).expand(int(vis.sum()), 2), atol=1e-4)
    assert torch.equal(prep.radii[vis], pre.radii[vis]) and torch.allclose(prep.conic[vis], pre.conic[vis])
    # near plane: a Gaussian at view depth 0.1 is culled at 0.2 and kept at 0.01
    fwd = cam.viewmatrix[:3, 2]                     # world-space direction of +z in view space (transposed matrix: column 2)
    near = (cam.campos + 0.1 * fwd)[None]
    kw = dict(scales=torch.full((1, 3), 0.01), rotations=torch.tensor([[1.0, 0, 0, 0]]), opacities=torch.tensor([[0.5]]))
    p_far = O.preprocess(near, kw["scales"], kw["rotations"], kw["opacities"], None, torch.ones(1, 3), None, S0)
    p_near = O.preprocess(near, kw["scales"], kw["rotations"], kw["opacities"], None, torch.ones(1, 3), None,
                          O.OracleSettings(**base, near_z=0.01))
    assert int(p_far.radii[0]) == 0 and int(p_near.radii[0]) > 0 and abs(float(p_near.depth[0]) - 0.1) < 1e-5
    # alpha clamp: an opaque splat's peak alpha is the clamp itself
    one = O.preprocess(torch.zeros(1, 3), torch.full((1, 3), 0.3), kw["rotations"], torch.tensor([[1.0]]), None, torch.ones(1, 3), None, S0)
    b1 = O.bin_and_sort(one, S0)
    lo = O.render_tiles(one, b1, S0).alpha.max()
    hi = O.render_tiles(one, b1, O.OracleSettings(**base, alpha_max=0.999)).alpha.max()
    assert abs(float(lo) - 0.99) < 1e-3 and float(hi) > 0.995


def test_unblended_gaussians_have_exactly_zero_gradients():
    """The premise of k_preprocess_bwd's zero-row shortcut and of the contributor bytes of the multi-GPU gather, checked on
    the spec side: a Gaussian that no pixel blends (culled, off screen, or hidden behind the depth at which its tiles
    saturate) receives EXACTLY zero gradients on every parameter -- its whole screen-space row is zero, not just small."""
    torch.manual_seed(0)
    sc = synth.make_scene(400, 3, 0.05, 0.5, seed=7)             # large, mostly opaque splats: deep layers are hidden
    sc = sc._replace(opacities=torch.full_like(sc.opacities, 0.97))
    cam = synth.look_at_camera(64, 48, (0.2, 0.1, -3.0))
    tgt = torch.full((48, 64), 2.8)
    touch = dict(touch_depth=tgt, touch_weight=torch.ones(48, 64), depth_loss="l1", depth_loss_mult=0.2)
    out, ins = _run(sc, cam, torch.float64, touch=touch)
    out.pre.rgb.retain_grad()
    g = torch.Generator().manual_seed(5)
    grgb = torch.rand(3, 48, 64, generator=g).double() + 0.1      # strictly positive: every blended splat gets colour gradient
    ((out.color * grgb).sum() + (out.depth * 0.3).sum() + (out.alpha * 0.2).sum() + out.touch_loss).backward()
    blended = (out.pre.rgb.grad != 0).any(1)
    visible = out.radii > 0
    assert int(blended.sum()) > 20 and int((visible & ~blended).sum()) > 20, "scene must have both blended and hidden splats"
    assert not bool((blended & ~visible).any())
    for name, t in zip(("means3D", "scales", "rotations", "opacities", "shs"), ins):
        gr = t.grad.reshape(t.shape[0], -1)
        assert bool((gr[~blended] == 0).all()), f"{name}: a Gaussian no pixel blended has a non-zero gradient"
        assert bool((gr[blended] != 0).any(1).all()), f"{name}: a blended Gaussian with an all-zero gradient"
