"""Small end-to-end case for compute-sanitizer: operator forward + fused-touch backward (whole image, a tile-row band
with the emulated peer gather, speculative sizing with an overflowing hint), one train step with refine, the
gsplat-style surface.  Sizes are tiny: the sanitizer slows kernels down 10-100x."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import touchgs_b200 as T  # noqa: E402

dev = torch.device("cuda:0")
synth = T.synth
sc = synth.make_scene(6000, 3, 0.01, 0.15, seed=5)
cam = synth.look_at_camera(192, 176, (0.3, 0.2, -2.8))
rs = T.GaussianRasterizationSettings(176, 192, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0, cam.viewmatrix.to(dev),
                                     cam.projmatrix.to(dev), 3, cam.campos.to(dev), False, False)
P = {k: getattr(sc, k).to(dev) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
g = torch.Generator().manual_seed(0)
gt = torch.rand(3, 176, 192, generator=g).to(dev)
with torch.no_grad():
    d = T.GaussianRasterizer(rs)(P["means3D"], None, P["opacities"], shs=P["shs"], scales=P["scales"], rotations=P["rotations"])[2]
tgt, wgt = synth.make_touch_maps(d[0].cpu() + 0.02, seed=0, n_patches=3, patch_radius=12)
tgt, wgt = tgt.to(dev), wgt.to(dev)


def step(**kw):
    ins = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    ras = T.GaussianRasterizer(rs)
    out = ras(ins["means3D"], None, ins["opacities"], shs=ins["shs"], scales=ins["scales"], rotations=ins["rotations"],
              touch_depth=tgt, touch_weight=wgt, depth_loss="l1", depth_loss_mult=0.2, **kw)
    T.photometric_loss(out[0], gt, 0.2).backward()
    torch.cuda.synchronize()
    return ras.last_num_rendered


I = step()
step(rendered_hint=int(I * 0.4))          # overflowing speculation: clamped launch, then the exact re-run
step(rendered_hint=int(I * 1.3))
step(tile_rows=(2, 6))                    # a band of a tile-row shard (8-warp scatter units, band-aware SH staging)
for binding in ("ctypes", "ext"):
    T._lib.use_binding(binding)
    step()
# the multi-GPU gather, one device playing 2 and 8 ranks, with the contributor bytes behind the rows and without
# (k_render_bwd<FLAGS>, k_preprocess_bwd<STAGE, GATHER, SKIP = 0 | 2>); the same function the GPU test suite runs
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_fullsize_parity as tfp  # noqa: E402
for world, flags in ((2, True), (8, True), (2, False)):
    tfp.test_emulated_peer_gather_is_bit_identical_to_summed_buffers(world, flags)
torch.cuda.synchronize()
# trainer step with refine (activations, SSIM loss, Adam, densify)
raw = [sc.means3D, sc.shs, torch.logit(sc.opacities.reshape(-1).clamp(1e-4, 1 - 1e-4)), torch.log(sc.scales), sc.rotations]
tr = T.TouchGSTrainer(*[t.to(dev) for t in raw], T.TrainConfig(sh_degree=3, refine_every=2, warmup_length=0, sh_degree_interval=0,
                                                              densify_grad_thresh=1e-7))
for _ in range(3):
    tr.train_step(rs, gt, tgt, wgt)
torch.cuda.synchronize()
# input pipeline kernels
T.dataset.decode_touch_maps((tgt * 1000).to(torch.int16).cpu().numpy().view("uint16"), None, 1e-3, "SIMPLE_LOSS", device="cuda")
torch.cuda.synchronize()
print("sanitize_case done:", I, tr.num_points)
