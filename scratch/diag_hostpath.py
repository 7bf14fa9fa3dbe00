"""Host-path cost of one operator step: tiny scene (GPU work negligible) so wall time per step == host overhead.
Prints per-phase perf_counter means and a cProfile top list."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import touchgs_b200 as T

dev = torch.device("cuda:0")
N, H, W = int(os.environ.get("DN", 2000)), 1080, 1920
sc = T.synth.make_scene(N, 3, 0.002, 0.02, 0)
cam = T.synth.orbit_cameras(W, H, 1, 3.0, 0)[0]
p = {k: getattr(sc, k).to(dev).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
bg = torch.zeros(3, device=dev)
view, proj, campos = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev)
gt = torch.rand(3, H, W, device=dev)
tgt = torch.full((H, W), 3.0, device=dev); wgt = torch.ones(H, W, device=dev)
inv = 1.0 / (3 * H * W)
hint = [0]
ph = [0.0, 0.0, 0.0]

def step():
    t0 = time.perf_counter()
    rs = T.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, view, proj, 3, campos, False, False)
    for v in p.values():
        v.grad = None
    ras = T.GaussianRasterizer(rs)
    color, radii, depth, alpha, resid = ras(p["means3D"], None, p["opacities"], shs=p["shs"], scales=p["scales"],
                                            rotations=p["rotations"], touch_depth=tgt, touch_weight=wgt, depth_loss="l1",
                                            depth_loss_mult=0.2, rendered_hint=hint[0])
    hint[0] = int(ras.last_num_rendered * 1.05) + 4096
    t1 = time.perf_counter()
    loss = (color - gt).abs().sum() * inv
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    ph[0] += t1 - t0; ph[1] += t2 - t1; ph[2] += t3 - t2

for _ in range(20):
    step()
torch.cuda.synchronize()
ph[:] = [0, 0, 0]
n = 200
t0 = time.perf_counter()
for _ in range(n):
    step()
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"N={N}: {1e6*(t1-t0)/n:.1f} us/step wall; host phases us: fwd {1e6*ph[0]/n:.1f} loss {1e6*ph[1]/n:.1f} bwd {1e6*ph[2]/n:.1f}")
pr = cProfile.Profile(); pr.enable()
for _ in range(100):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
