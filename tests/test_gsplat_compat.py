"""gsplat-0.1-style three-call surface (touch-gs_b200/gsplat_compat.py) against the CPU oracle run with
the same convention switches (near plane = clip_thresh, principal point, alpha clamp 0.999, pixel-centre sampling).
Integer results bit-exact, floats 1e-4 relative (north star)."""
import pytest
import torch

from helpers import O, T, synth, rel_inf, assert_close_tensor

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
G = None


@pytest.fixture(scope="module", autouse=True)
def _require_cuda(tgs_lib):
    global G
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from importlib import import_module
    G = import_module("touch-gs_b200.gsplat_compat")


def _setup(N=2500, W=176, H=130, seed=5, cx_off=7.25, cy_off=-4.5):
    sc = synth.make_scene(N, 0, 0.02, 0.25, seed=seed)
    cam = synth.look_at_camera(W, H, (0.3, 0.2, -2.6))
    fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
    cx, cy = W / 2.0 + cx_off, H / 2.0 + cy_off
    g = torch.Generator().manual_seed(seed)
    colors = torch.rand(N, 3, generator=g)
    quats = sc.rotations * (0.5 + torch.rand(N, 1, generator=g))          # NOT normalised: the op normalises inside
    bg = torch.tensor([0.2, 0.5, 0.1])
    S = O.OracleSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.3, cam.viewmatrix, cam.projmatrix, 0, cam.campos,
                         alpha_max=0.999, near_z=0.01, principal=(cx_off, cy_off), pixel_offset=0.5)
    return sc, cam, (fx, fy, cx, cy), colors, quats, bg, S


def test_project_gaussians_matches_oracle():
    sc, cam, (fx, fy, cx, cy), colors, quats, bg, S = _setup()
    H, W = S.image_height, S.image_width
    rot = quats / quats.norm(dim=-1, keepdim=True)
    pre = O.preprocess(sc.means3D, sc.scales, rot, sc.opacities, None, colors, None, S)
    view, proj = cam.viewmatrix.t().contiguous(), cam.projmatrix.t().contiguous()      # column-vector convention
    xys, depths, radii, conics, nth, cov3d = G.project_gaussians(
        sc.means3D.to(DEV), sc.scales.to(DEV), 1.3, quats.to(DEV), view.to(DEV), proj.to(DEV), fx, fy, cx, cy, H, W,
        clip_thresh=0.01)
    vis = pre.radii > 0
    assert int(vis.sum()) > 500
    assert torch.equal(radii.cpu(), pre.radii) and torch.equal(nth.cpu(), pre.tiles_touched.int())
    assert_close_tensor(xys.cpu()[vis], pre.xy[vis], "xys", 1e-6)
    assert_close_tensor(depths.cpu()[vis], pre.depth[vis], "depths", 1e-6)
    assert_close_tensor(conics.cpu()[vis], pre.conic[vis], "conics", 1e-5)
    # near plane really is clip_thresh: something between 0.01 and 0.2 of the camera survives
    close = synth.make_scene(200, 0, 0.01, 0.02, seed=1)
    m = close.means3D * 0.02 + cam.campos + torch.tensor([0.0, 0.0, 0.1]) @ cam.viewmatrix[:3, :3].t()
    _, d2, r2, *_ = G.project_gaussians(m.to(DEV), close.scales.to(DEV), 1.0, close.rotations.to(DEV), view.to(DEV),
                                        proj.to(DEV), fx, fy, W / 2.0, H / 2.0, H, W, clip_thresh=0.01)
    sel = r2 > 0
    assert bool(sel.any()) and float(d2[sel].max()) < 0.2


def test_rasterize_and_gradients_match_oracle():
    sc, cam, (fx, fy, cx, cy), colors, quats, bg, S = _setup()
    H, W = S.image_height, S.image_width
    g = torch.Generator().manual_seed(9)
    Gi, Ga = torch.rand(H, W, 3, generator=g) / (H * W), torch.rand(H, W, generator=g) / (H * W)
    # oracle
    ins = [t.clone().requires_grad_(True) for t in (sc.means3D, sc.scales, quats, colors, sc.opacities)]
    rot = ins[2] / ins[2].norm(dim=-1, keepdim=True)
    pre = O.preprocess(ins[0], ins[1], rot, ins[4], None, ins[3], None, S)
    bins = O.bin_and_sort(pre, S)
    img = O.render_tiles(pre, bins, S)
    ((img.color.permute(1, 2, 0) * Gi).sum() + (img.alpha * Ga).sum()).backward()
    # ours: project -> rasterize
    view, proj = cam.viewmatrix.t().contiguous().to(DEV), cam.projmatrix.t().contiguous().to(DEV)
    cin = [t.to(DEV).clone().requires_grad_(True) for t in (sc.means3D, sc.scales, quats, colors, sc.opacities)]
    xys, depths, radii, conics, nth, _ = G.project_gaussians(cin[0], cin[1], 1.3, cin[2], view, proj, fx, fy, cx, cy, H, W,
                                                             clip_thresh=0.01)
    out, alpha = G.rasterize_gaussians(xys, depths, radii, conics, nth, cin[3], cin[4], H, W, bg.to(DEV), return_alpha=True)
    assert out.shape == (H, W, 3) and alpha.shape == (H, W)
    ((out * Gi.to(DEV)).sum() + (alpha * Ga.to(DEV)).sum()).backward()
    assert_close_tensor(out, img.color.permute(1, 2, 0), "image", 1e-4, 5e-4)
    assert_close_tensor(alpha, img.alpha, "alpha", 1e-4, 5e-4)
    for nm, a, b in zip(("means3d", "scales", "quats", "colors", "opacity"), cin, ins):
        assert_close_tensor(a.grad, b.grad, "v_" + nm, 1e-4)


def test_depth_as_colour_and_single_channel():
    """The fork renders depth by a second rasterize call with depth as the colour: C = 1 works and the
    default background is ones."""
    sc, cam, (fx, fy, cx, cy), colors, quats, bg, S = _setup(N=800)
    H, W = S.image_height, S.image_width
    view, proj = cam.viewmatrix.t().contiguous().to(DEV), cam.projmatrix.t().contiguous().to(DEV)
    xys, depths, radii, conics, nth, _ = G.project_gaussians(sc.means3D.to(DEV), sc.scales.to(DEV), 1.3, quats.to(DEV), view, proj,
                                                             fx, fy, cx, cy, H, W)
    d_img, alpha = G.rasterize_gaussians(xys, depths, radii, conics, nth, depths[:, None], sc.opacities.to(DEV), H, W,
                                         torch.zeros(1, device=DEV), return_alpha=True)
    rgb = G.rasterize_gaussians(xys, depths, radii, conics, nth, colors.to(DEV), sc.opacities.to(DEV), H, W)
    assert d_img.shape == (H, W, 1) and rgb.shape == (H, W, 3)
    hit = alpha > 0.5
    exp_d = d_img[..., 0][hit] / alpha[hit]
    assert float(exp_d.min()) > 0.01 and float(exp_d.max()) < 10.0
    # default background = ones: an empty pixel is white
    empty = alpha == 0
    if bool(empty.any()):
        assert float((rgb[empty] - 1.0).abs().max()) == 0.0


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_spherical_harmonics_matches_oracle(deg):
    g = torch.Generator().manual_seed(deg)
    N, K = 3000, 16
    dirs = torch.randn(N, 3, generator=g) * 2.0
    coeffs = torch.randn(N, K, 3, generator=g) * 0.05
    nb = (deg + 1) ** 2
    ref_in = coeffs[:, :nb].clone().requires_grad_(True)
    rgb, clamped = O.sh_to_rgb(deg, dirs / dirs.norm(dim=-1, keepdim=True), ref_in)      # = raw + 0.5, clamped at 0
    assert not bool(clamped.any())
    w = torch.randn(N, 3, generator=g)
    (rgb * w).sum().backward()
    c = coeffs.to(DEV).requires_grad_(True)
    out = G.spherical_harmonics(deg, dirs.to(DEV), c)
    (out * w.to(DEV)).sum().backward()
    assert_close_tensor(out + 0.5, rgb, "sh colours", 1e-5)
    assert_close_tensor(c.grad[:, :nb], ref_in.grad, "v_coeffs", 1e-5)
    assert float(c.grad[:, nb:].abs().sum()) == 0.0
