"""The "reference-structure CUDA" arm (csrc/refstructure.cu; BASELINE.md §3 column 2) against the CPU oracle at
small sizes, and the PRODUCT kernels against it at BASELINE's full size c3 (1M Gaussians, 1080p), where the CPU
oracle is too slow: two independently structured traversals (one pixel per thread, no culling, gathered data,
per-thread atomics, unfused PyTorch touch loss  vs.  TMA-staged packed lists, exact warp cull, four pixels per
thread, fused touch gradient) must agree BIT-EXACTLY on every integer result (sorted keys / ids, ranges,
n_contrib) and on the forward images, and within 1e-4 relative on every gradient tensor."""
import pytest
import torch

from helpers import O, T, synth, oracle_settings, cuda_settings, rel_inf, assert_close_tensor

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
R = T.refstructure


@pytest.fixture(scope="module", autouse=True)
def _require_cuda(tgs_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"


def _touch(depth_img, seed):
    return synth.make_touch_maps(depth_img + 0.02, seed=seed, n_patches=3, patch_radius=12)


@pytest.mark.parametrize("case", [dict(N=1000, W=128, H=128, deg=0, smin=0.02, smax=0.2, eye=(0.5, 0.3, -3.0), seed=0),
                                  dict(N=3000, W=203, H=117, deg=2, smin=0.02, smax=0.3, eye=(0.2, -0.4, -2.2), seed=2)],
                         ids=["c1", "ragged"])
def test_refstructure_matches_oracle(case):
    c = case
    sc = synth.make_scene(c["N"], c["deg"], c["smin"], c["smax"], seed=c["seed"])
    cam = synth.look_at_camera(c["W"], c["H"], c["eye"])
    bg = (0.1, 0.2, 0.3)
    S = oracle_settings(cam, c["deg"], bg=bg)
    rs = cuda_settings(cam, c["deg"], DEV, bg=bg)
    pre = O.preprocess(sc.means3D, sc.scales, sc.rotations, sc.opacities, sc.shs, None, None, S)
    bins = O.bin_and_sort(pre, S)
    dv = [t.to(DEV) for t in (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations)]
    st = R.forward_state(*dv, rs)
    # integer stage: BIT-EXACT against the oracle (the upstream single-sort formulation of A2-A4)
    assert st["num_rendered"] == int(bins.keys.numel())
    assert torch.equal(st["keys"].cpu(), bins.keys) and torch.equal(st["vals"].cpu().long(), bins.vals.long())
    assert torch.equal(st["ranges"].cpu().long(), bins.ranges.long())
    # float stage + gradients with the UNFUSED PyTorch touch loss against the oracle's touch loss
    g = torch.Generator().manual_seed(1)
    grgb = torch.rand(3, c["H"], c["W"], generator=g) / (3 * c["H"] * c["W"])
    base = O.rasterize(sc.means3D, sc.opacities, S, shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
    tgt, wgt = _touch(base.depth[0], c["seed"])
    ins = [t.clone().requires_grad_(True) for t in (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations)]
    ref = O.rasterize(ins[0], ins[1], S, shs=ins[2], scales=ins[3], rotations=ins[4], touch_depth=tgt,
                      touch_weight=wgt, depth_loss="l1", depth_loss_mult=0.2)
    ((ref.color * grgb).sum() + ref.touch_loss).backward()
    cin = [t.clone().requires_grad_(True) for t in dv]
    color, radii, draw, alpha = R.rasterize_refstructure(*cin, rs)
    loss = (color * grgb.to(DEV)).sum() + R.touch_depth_loss_unfused(draw, alpha, tgt.to(DEV), wgt.to(DEV), 0.2, "l1")
    loss.backward()
    assert torch.equal(radii.cpu(), ref.radii)
    assert_close_tensor(color, ref.color, "color", 1e-4, 5e-4)
    assert_close_tensor(alpha, ref.alpha, "alpha", 1e-4, 5e-4)
    dhat = torch.where(alpha > 0, draw / alpha.clamp_min(1e-30), torch.zeros_like(draw))
    assert_close_tensor(dhat, ref.depth, "depth", 1e-4, 5e-4)
    for name, a, b in zip(("means3D", "opacities", "shs", "scales", "rotations"), cin, ins):
        assert_close_tensor(a.grad, b.grad, "d" + name, 1e-4)


def test_product_kernels_match_refstructure_at_full_size_c3():
    cfg = synth.CONFIGS["c3"]
    H, W, deg = cfg["H"], cfg["W"], cfg["sh_degree"]
    sc = synth.make_scene(cfg["N"], deg, cfg["smin"], cfg["smax"], seed=0)
    cam = synth.orbit_cameras(W, H, 8, 3.0, 0)[0]
    rs = cuda_settings(cam, deg, DEV)
    dv = [t.to(DEV) for t in (sc.means3D, sc.opacities, sc.shs, sc.scales, sc.rotations)]
    m, o, sh, s, r = dv
    ours = T.inspect_state.forward_state(m, o, rs, shs=sh, scales=s, rotations=r,
                                         opt=T.TouchOptions(depth_normalize=False))
    ref = R.forward_state(m, o, sh, s, r, rs)
    assert ours["num_rendered"] == ref["num_rendered"] > 1_000_000
    # two-phase (depth, then 16-bit tile) sort == single 45-bit (tile | depth) sort, entry by entry
    assert torch.equal(ours["keys"], ref["keys"]), "sorted 64-bit keys differ"
    assert torch.equal(ours["vals"], ref["vals"]), "sorted Gaussian ids differ"
    assert torch.equal(ours["ranges"], ref["ranges"])
    # exact warp-level cull + shared per-pair arithmetic: the same pairs blend, in the same order
    assert torch.equal(ours["n_contrib"], ref["n_contrib"]), "n_contrib differs: the cull is not exact"
    assert torch.equal(ours["final_T"], ref["final_T"])
    assert torch.equal(ours["alpha"], ref["alpha"])
    assert rel_inf(ours["color"], ref["color"]) < 1e-6 and rel_inf(ours["depth"], ref["depth_raw"]) < 1e-6
    del ours, ref
    # gradients: fused touch depth-L1 (ours) vs PyTorch loss on the rendered images (reference structure)
    with torch.no_grad():
        _, _, d, _, _ = T.GaussianRasterizer(rs)(m, None, o, shs=sh, scales=s, rotations=r)
    tgt, wgt = synth.make_touch_maps(d[0].cpu() + 0.01, seed=0)
    tgt, wgt = tgt.to(DEV), wgt.to(DEV)
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(3, H, W, generator=g).to(DEV)
    a_in = [t.clone().requires_grad_(True) for t in dv]
    color, _, _, _, _ = T.GaussianRasterizer(rs)(a_in[0], None, a_in[1], shs=a_in[2], scales=a_in[3], rotations=a_in[4],
                                                 touch_depth=tgt, touch_weight=wgt, depth_loss="l1", depth_loss_mult=0.2)
    (color - gt).abs().mean().backward()
    b_in = [t.clone().requires_grad_(True) for t in dv]
    color_b, _, draw, alpha = R.rasterize_refstructure(*b_in, rs)
    ((color_b - gt).abs().mean() + R.touch_depth_loss_unfused(draw, alpha, tgt, wgt, 0.2, "l1")).backward()
    for name, a, b in zip(("means3D", "opacities", "shs", "scales", "rotations"), a_in, b_in):
        assert torch.isfinite(a.grad).all()
        assert_close_tensor(a.grad, b.grad, "d" + name, 1e-4)
