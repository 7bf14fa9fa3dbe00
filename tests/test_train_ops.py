"""GPU parity of the train-step kernels (csrc/train_ops.cu; SURVEY §8f N1, BASELINE config c5) against
oracle/train_oracle.py (torch autograd / torch.optim.Adam on the CPU).  Tolerances: 1e-4 relative per tensor for
losses and gradients (north star), 1e-6 for Adam (same arithmetic, different rounding of one lerp), bit-exact for
the integer results of the refine step (counts, offsets, source ids) and for every copied value."""
import math

import pytest
import torch

from helpers import O, T, synth, oracle_settings, cuda_settings, rel_inf, assert_close_tensor

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
TO = O.train_oracle


@pytest.fixture(scope="module", autouse=True)
def _require_cuda(tgs_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"


@pytest.mark.parametrize("shape", [(128, 128), (117, 203), (16, 16), (40, 9)])
@pytest.mark.parametrize("lam", [0.2, 0.0, 1.0])
def test_photometric_loss_matches_oracle(shape, lam):
    H, W = shape
    g = torch.Generator().manual_seed(H * 1000 + W)
    a = torch.rand(3, H, W, generator=g)
    b = (a + 0.2 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    x = a.clone().requires_grad_(True)
    ref = TO.photometric_loss(x, b, lam)
    (ref * 1.7).backward()
    xc = a.to(DEV).requires_grad_(True)
    out = T.photometric_loss(xc, b.to(DEV), lam)
    (out * 1.7).backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * max(abs(float(ref)), 1e-3)
    assert_close_tensor(xc.grad, x.grad, "dloss/dcolor", 1e-4)


def test_photometric_loss_band_partition_with_halo():
    H, W = 96, 70
    g = torch.Generator().manual_seed(5)
    a, b = torch.rand(3, H, W, generator=g), torch.rand(3, H, W, generator=g)
    x = a.clone().requires_grad_(True)
    TO.photometric_loss(x, b, 0.2).backward()
    total, grad = 0.0, torch.zeros(3, H, W, device=DEV)
    for (y0, y1) in ((0, 32), (32, 64), (64, 96)):
        xc = a.to(DEV).requires_grad_(True)
        o0, o1 = max(0, y0 - 16), min(H, y1 + 16)                  # band + one tile row of halo
        l = T.photometric_loss(xc, b.to(DEV), 0.2, rows=(y0, y1), out_rows=(o0, o1))
        l.backward()
        ref_band = TO.photometric_loss(a, b, 0.2, rows=(y0, y1))
        assert abs(float(l) - float(ref_band)) <= 1e-5 * abs(float(ref_band))
        assert float(xc.grad[:, :o0].abs().sum()) == 0.0 and float(xc.grad[:, o1:].abs().sum()) == 0.0
        total += float(l)
        grad += xc.grad
    assert abs(total - float(TO.photometric_loss(a, b, 0.2))) < 1e-5
    assert_close_tensor(grad, x.grad, "sum of band gradients", 1e-4)


def test_activate_matches_oracle():
    g = torch.Generator().manual_seed(3)
    N = 1237
    s, q, o = torch.randn(N, 3, generator=g) - 4, torch.randn(N, 4, generator=g), torch.randn(N, 1, generator=g) * 3
    ws, wq, wo = torch.randn(N, 3, generator=g), torch.randn(N, 4, generator=g), torch.randn(N, 1, generator=g)
    ins = [t.clone().requires_grad_(True) for t in (s, q, o)]
    rs_, rq, ro = TO.activate(*ins)
    ((rs_ * ws).sum() + (rq * wq).sum() + (ro * wo).sum()).backward()
    cin = [t.to(DEV).requires_grad_(True) for t in (s, q, o)]
    cs, cq, co = T.activate(*cin)
    ((cs * ws.to(DEV)).sum() + (cq * wq.to(DEV)).sum() + (co * wo.to(DEV)).sum()).backward()
    for nm, a, b in (("scales", cs, rs_), ("rot", cq, rq), ("opac", co, ro)):
        assert_close_tensor(a, b, nm, 1e-6)
    for nm, a, b in zip(("dscales_log", "dquats", "dopacity_logit"), cin, ins):
        assert_close_tensor(a.grad, b.grad, nm, 1e-5)


def test_adam_matches_torch_optim():
    g = torch.Generator().manual_seed(4)
    N, K = 1001, 16                                       # 1001: numel % 4 != 0 exercises the scalar tail
    shapes = dict(means=(N, 3), dc=(N, 1, 3), rest=(N, K - 1, 3), opacity=(N,), scales=(N, 3), quats=(N, 4))
    lrs = dict(means=1.6e-4, dc=2.5e-3, rest=2.5e-3 / 20, opacity=5e-2, scales=5e-3, quats=1e-3)
    p = {k: torch.randn(*s, generator=g) for k, s in shapes.items()}
    gr = {k: torch.randn(*s, generator=g) * (10.0 ** torch.randint(-6, 1, s, generator=g).float()) for k, s in shapes.items()}
    names = list(shapes)
    steps = 3
    rp, rm, rv = TO.adam_reference([p[k] for k in names], [gr[k] for k in names], [lrs[k] for k in names], steps)
    ref = {k: (rp[i], rm[i], rv[i]) for i, k in enumerate(names)}
    # ours: ONE SH tensor [N,K,3] with the two learning rates selected by (index % 3K) < 3
    dev = {k: v.to(DEV) for k, v in p.items()}
    dg = {k: v.to(DEV) for k, v in gr.items()}
    dev["shs"] = torch.cat([dev.pop("dc"), dev.pop("rest")], 1).contiguous()
    dg["shs"] = torch.cat([dg.pop("dc"), dg.pop("rest")], 1).contiguous()
    m = {k: torch.zeros_like(v) for k, v in dev.items()}
    v = {k: torch.zeros_like(x) for k, x in dev.items()}
    groups = []
    for k in dev:
        d = dict(param=dev[k], grad=dg[k], exp_avg=m[k], exp_avg_sq=v[k], lr=lrs.get(k, lrs["dc"]))
        if k == "shs":
            d.update(lr_tail=lrs["rest"], period=3 * K, head=3)
        groups.append(d)
    own0, _ = T._lib.launch_counts()
    for t in range(1, steps + 1):
        T.adam_step(groups, t, (0.9, 0.999), 1e-15)
    torch.cuda.synchronize()
    assert T._lib.launch_counts()[0] - own0 == steps, "one launch per step for all groups"
    for k in ("means", "opacity", "scales", "quats"):
        for a, b, nm in zip((dev[k], m[k], v[k]), ref[k], ("param", "exp_avg", "exp_avg_sq")):
            assert rel_inf(a, b) < 2e-6, (k, nm, rel_inf(a, b))
    for a, b0, b1, nm in zip((dev["shs"], m["shs"], v["shs"]), ref["dc"], ref["rest"], ("param", "exp_avg", "exp_avg_sq")):
        assert rel_inf(a[:, :1], b0) < 2e-6 and rel_inf(a[:, 1:], b1) < 2e-6, nm


@pytest.mark.parametrize("allow,screen", [(True, False), (False, False), (True, True)])
def test_densify_matches_oracle(allow, screen):
    g = torch.Generator().manual_seed(6)
    N, K = 3001, 16
    means, shs = torch.randn(N, 3, generator=g), torch.randn(N, K, 3, generator=g)
    op = torch.randn(N, generator=g) * 2
    sl = torch.log(torch.rand(N, 3, generator=g) * 0.03 + 0.001)
    sl[:7] = math.log(0.8)
    q = torch.randn(N, 4, generator=g)
    acc = torch.rand(N, generator=g) * 8e-4
    vc = torch.randint(0, 4, (N,), generator=g).int()
    noise = torch.randn(N, 2, 3, generator=g)
    # screen-size rules (largest screen radius since the last refine): split above 40 px, cull above 90 px
    mr = torch.randint(0, 120, (N,), generator=g).int() if screen else None
    ocfg = TO.DensifyConfig(split_screen_radius=40.0, cull_screen_radius=90.0) if screen else TO.DensifyConfig()
    tcfg = T.TrainConfig(split_screen_radius=40.0, cull_screen_radius=90.0) if screen else T.TrainConfig()
    ref = TO.densify_reference(means, shs, op, sl, q, acc, vc, noise, ocfg, allow_split_dup=allow, max_radii=mr)
    params = dict(means=means, shs=shs, opacity_logit=op, scales_log=sl, quats=q)
    params = {k: v.to(DEV) for k, v in params.items()}
    m = {k: torch.randn(v.shape, generator=g).to(DEV) for k, v in params.items()}
    v = {k: torch.rand(x.shape, generator=g).to(DEV) for k, x in params.items()}
    np_, nm, nv, src = T.densify(params, m, v, acc.to(DEV), vc.to(DEV), noise.to(DEV), tcfg, allow,
                                 None if mr is None else mr.to(DEV))
    M = ref["means"].shape[0]
    assert np_["means"].shape[0] == M
    rsrc, rnew = ref["src"], ref["is_new"]
    want = torch.where(rnew, -(rsrc + 1), rsrc).int()
    assert torch.equal(src.cpu(), want), "source ids / new flags differ"
    kept = ~rnew
    for k, rk in (("means", "means"), ("shs", "shs"), ("opacity_logit", "opacity_logit"), ("scales_log", "scales_log"), ("quats", "quats")):
        a, b = np_[k].cpu(), ref[rk]
        assert torch.equal(a[kept], b[kept]), f"{k}: carried-over values must be bit-identical"
        assert_close_tensor(a, b, k, 1e-6)
        # Adam moments travel with carried-over Gaussians and are zero for new ones
        assert torch.equal(nm[k].cpu()[kept], m[k].cpu()[rsrc[kept]]) and torch.equal(nv[k].cpu()[kept], v[k].cpu()[rsrc[kept]])
        assert float(nm[k].cpu()[rnew].abs().sum()) == 0.0 and float(nv[k].cpu()[rnew].abs().sum()) == 0.0


def test_densify_everything_culled_and_empty_population():
    """refine may cull every Gaussian (M == 0); the next plan / step on the empty population must not raise."""
    N, K = 64, 4
    params = dict(means=torch.zeros(N, 3), shs=torch.zeros(N, K, 3), opacity_logit=torch.full((N,), -9.0),
                  scales_log=torch.full((N, 3), -4.0), quats=torch.ones(N, 4))
    params = {k: v.to(DEV) for k, v in params.items()}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(x) for k, x in params.items()}
    z = torch.zeros(N, device=DEV)
    p2, m2, v2, src = T.densify(params, m, v, z, z.int(), torch.zeros(N, 2, 3, device=DEV), T.TrainConfig(), True)
    assert p2["means"].shape[0] == 0 and src.numel() == 0
    e = torch.zeros(0, device=DEV)
    p3, _, _, src3 = T.densify(p2, m2, v2, e, e.int(), torch.zeros(0, 2, 3, device=DEV), T.TrainConfig(), True)
    assert p3["means"].shape[0] == 0 and src3.numel() == 0


def test_densify_stats_matches_oracle():
    g = torch.Generator().manual_seed(7)
    N = 5000
    d2 = torch.randn(N, 3, generator=g)
    radii = torch.randint(-1, 40, (N,), generator=g).int()
    acc, vc, mr = torch.rand(N, generator=g), torch.randint(0, 5, (N,), generator=g).int(), torch.randint(0, 30, (N,), generator=g).int()
    r = TO.densify_stats_reference(d2, radii, acc, vc, mr)
    a, c, m = acc.to(DEV), vc.to(DEV), mr.to(DEV)
    lib = T._lib.load()
    import ctypes as C
    P = lambda t: C.c_void_p(t.data_ptr())
    d2d, rd = d2.to(DEV), radii.to(DEV)
    T._lib.check(lib.tgs_densify_stats(N, P(d2d), P(rd), P(a), P(c), P(m), None), "tgs_densify_stats")
    torch.cuda.synchronize()
    assert rel_inf(a, r[0]) < 1e-6 and torch.equal(c.cpu(), r[1]) and torch.equal(m.cpu(), r[2])


def _oracle_train_grads(sc, cam, deg, gt, tgt, wgt, cfg):
    """One oracle forward/backward of the full loss on the RAW parameters."""
    S = oracle_settings(cam, deg)
    raw = dict(means=sc.means3D, shs=sc.shs, opacity_logit=torch.logit(sc.opacities.reshape(-1).clamp(1e-4, 1 - 1e-4)),
               scales_log=torch.log(sc.scales), quats=sc.rotations * 1.7)
    ins = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    scales, rot, opac = TO.activate(ins["scales_log"], ins["quats"], ins["opacity_logit"])
    out = O.rasterize(ins["means"], opac, S, shs=ins["shs"], scales=scales, rotations=rot, touch_depth=tgt, touch_weight=wgt,
                      depth_loss="l1", depth_loss_mult=cfg.depth_loss_mult)
    loss = TO.photometric_loss(out.color, gt, cfg.ssim_lambda) + out.touch_loss
    loss.backward()
    return raw, {k: v.grad for k, v in ins.items()}, out


def test_train_step_matches_oracle_and_adam_moves_parameters():
    c = dict(N=1500, W=128, H=112, deg=2, smin=0.02, smax=0.2, eye=(0.4, 0.3, -3.0), seed=11)
    sc = synth.make_scene(c["N"], c["deg"], c["smin"], c["smax"], seed=c["seed"])
    cam = synth.look_at_camera(c["W"], c["H"], c["eye"])
    g = torch.Generator().manual_seed(8)
    gt = torch.rand(3, c["H"], c["W"], generator=g)
    base = O.rasterize(sc.means3D, sc.opacities, oracle_settings(cam, c["deg"]), shs=sc.shs, scales=sc.scales, rotations=sc.rotations)
    tgt, wgt = synth.make_touch_maps(base.depth[0] + 0.02, seed=1, n_patches=3, patch_radius=12)
    cfg = T.TrainConfig(sh_degree=c["deg"], depth_loss_type="DEPTH_UNCERTAINTY_WEIGHTED_LOSS", uncertainty_weight=2.0,
                        refine_every=0, sh_degree_interval=0)
    raw, ref_g, ref_out = _oracle_train_grads(sc, cam, c["deg"], gt, tgt, wgt / 2.0, cfg)
    tr = T.TouchGSTrainer(*[raw[k].to(DEV) for k in ("means", "shs", "opacity_logit", "scales_log", "quats")], cfg)
    before = {k: v.clone() for k, v in tr.p.items()}
    rs = cuda_settings(cam, c["deg"], DEV)
    loss = tr.train_step(rs, gt.to(DEV), tgt.to(DEV), wgt.to(DEV))
    ref_photo = float(TO.photometric_loss(ref_out.color, gt, cfg.ssim_lambda))
    assert abs(float(loss) - ref_photo) <= 2e-4 * abs(ref_photo)
    for k in ("means", "shs", "opacity_logit", "scales_log", "quats"):
        assert_close_tensor(tr.last["grads"][k], ref_g[k], "d" + k, 1e-4)
    # first Adam step: every element with a non-negligible gradient moves by lr * sign(grad)
    lr = dict(means=cfg.lr_means, opacity_logit=cfg.lr_opacity, scales_log=cfg.lr_scales, quats=cfg.lr_quats)
    for k, l in lr.items():
        gk = ref_g[k].to(DEV)
        big = gk.abs() > 1e-3 * gk.abs().max()
        delta = (tr.p[k] - before[k])[big]
        assert torch.allclose(delta, -l * torch.sign(gk[big]), rtol=1e-3, atol=l * 1e-3), k
    # refine statistics of the step
    vis = ref_out.radii > 0
    assert torch.equal(tr.vis_count.cpu() > 0, vis)


def test_trainer_refine_changes_population_consistently():
    sc = synth.make_scene(4000, 1, 0.004, 0.06, seed=12)
    cam = synth.look_at_camera(160, 128, (0.3, 0.2, -3.0))
    raw = [sc.means3D, sc.shs, torch.logit(sc.opacities.reshape(-1).clamp(1e-4, 1 - 1e-4)), torch.log(sc.scales), sc.rotations]
    cfg = T.TrainConfig(sh_degree=1, refine_every=2, warmup_length=0, densify_grad_thresh=1e-7, reset_alpha_every=1,
                        sh_degree_interval=2)
    tr = T.TouchGSTrainer(*[t.to(DEV) for t in raw], cfg)
    rs = cuda_settings(cam, 1, DEV)
    gt = torch.rand(3, 128, 160, device=DEV)
    n0 = tr.num_points
    tr.train_step(rs, gt)
    assert tr.num_points == n0
    tr.train_step(rs, gt)                                  # step 2 -> refine + opacity reset
    n1 = tr.num_points
    assert n1 != n0, "refine should have changed the population"
    for d in (tr.p, tr.m, tr.v):
        assert all(t.shape[0] == n1 for t in d.values())
    assert tr.grad_accum.shape[0] == n1 and int(tr.vis_count.sum()) == 0
    assert float(torch.sigmoid(tr.p["opacity_logit"]).max()) <= 2 * cfg.cull_alpha_thresh + 1e-6
    # SH degree schedule (interval 2): steps 1 were degree 0 -> the higher bands got no gradient there; step 3 is degree 1
    l = tr.train_step(rs, gt)                              # the resized state keeps training
    assert math.isfinite(float(l))
    assert float(tr.last["grads"]["shs"][:, 1:].abs().sum()) > 0.0
    fresh = T.TouchGSTrainer(*[t.to(DEV) for t in raw], cfg)
    fresh.train_step(rs, gt)
    assert float(fresh.last["grads"]["shs"][:, 1:].abs().sum()) == 0.0 and float(fresh.last["grads"]["shs"][:, 0].abs().sum()) > 0.0
