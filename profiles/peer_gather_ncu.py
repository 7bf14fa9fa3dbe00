"""NVLink evidence for the fused exchange: ONE process drives two GPUs so that ncu can profile the
gather kernel.  cuda:0 plays rank 0 (renders tile rows [0, 34) into its [N,10] buffer), cuda:1 plays rank 1 (rows
[34, 68), its buffer lives in cuda:1's memory); tgs_backward_preprocess_gather then runs on cuda:0 and reads rank 1's
rows straight out of cuda:1's memory over NVLink (peer access enabled through torch).  c3 scene, camera 0.

    ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum -k regex:k_preprocess_bwd python profiles/peer_gather_ncu.py

FLAGS=1 (default): the buffers carry the contributor bytes (TgsSettings.contrib_flags) and the gather asks rank 1 only for the
rows it flagged; FLAGS=0: plain [N,10] rows, every row of a Gaussian whose span touches rank 1's band is read.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TGS_BINDING"] = "ctypes"
import torch  # noqa: E402
import touchgs_b200 as T  # noqa: E402
from importlib import import_module  # noqa: E402

R = import_module("touch-gs_b200.rasterizer")
L = T._lib
lib = L.load()
assert torch.cuda.device_count() >= 2, "needs two GPUs"
cfg = T.synth.CONFIGS["c3"]
H, W, deg, N = cfg["H"], cfg["W"], 3, cfg["N"]
sc = T.synth.make_scene(N, deg, cfg["smin"], cfg["smax"], 0)
cam = T.synth.orbit_cameras(W, H, 8, 3.0, 0)[0]
bands = T.sharding.even_bands(H, 2)
FLAGS = os.environ.get("FLAGS", "1") != "0"
NF = L.screen_grad_floats(N, FLAGS)
g = torch.Generator().manual_seed(0)
grgb_cpu = torch.rand(3, H, W, generator=g) / (3 * H * W)
state = []
for r in (0, 1):
    dev = torch.device("cuda", r)
    with torch.cuda.device(dev):
        rs = T.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, torch.zeros(3, device=dev), 1.0, cam.viewmatrix.to(dev),
                                             cam.projmatrix.to(dev), deg, cam.campos.to(dev), False, False)
        t = [x.to(dev).contiguous() for x in (sc.means3D, sc.opacities.reshape(-1), sc.shs, sc.scales, sc.rotations)]
        keep = []
        st, _ = R._make_settings(rs, T.TouchOptions(tile_rows=bands[r]), 16, keep)
        gs = R._make_gaussians(t[0], t[1], t[2], None, t[3], t[4], None)
        out = [torch.zeros(3, H, W, device=dev), torch.zeros(H, W, device=dev), torch.zeros(H, W, device=dev),
               torch.zeros(N, dtype=torch.int32, device=dev)]
        scratch = R._Scratch(dev)
        saved = L.TgsSaved()
        p = lambda x: C.c_void_p(x.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        L.check(lib.tgs_forward(C.byref(st), C.byref(gs), scratch.cb, None, p(out[0]), p(out[1]), p(out[2]), p(out[3]), None, None,
                                C.byref(saved), stream), "fwd")
        scratch.disarm()
        st.contrib_flags = 1 if FLAGS else 0
        sg = torch.zeros(NF, device=dev)
        grgb = grgb_cpu.to(dev)
        L.check(lib.tgs_backward_render(C.byref(st), C.byref(gs), C.byref(saved), p(grgb), None, None, None, None, p(sg), stream), "bwd")
        torch.cuda.synchronize(dev)
        state.append(dict(st=st, gs=gs, saved=saved, radii=out[3], sg=sg, keep=(keep, t, scratch.bufs, out, grgb, rs)))
# enable peer access cuda:0 -> cuda:1 explicitly (what symmetric memory / NCCL do for the multi-process path)
assert torch.cuda.can_device_access_peer(0, 1), "no P2P between cuda:0 and cuda:1"
_ = state[1]["sg"][:4].to("cuda:0")
cudart = C.CDLL("libcudart.so.12")
with torch.cuda.device(0):
    rc = cudart.cudaDeviceEnablePeerAccess(C.c_int(1), C.c_uint(0))
    assert rc in (0, 704), f"cudaDeviceEnablePeerAccess failed: {rc}"      # 704 = already enabled
    cudart.cudaGetLastError()
with torch.cuda.device(0):
    dev = torch.device("cuda:0")
    s0 = state[0]
    K = 16
    gr = dict(dmeans2D=torch.empty(N, 3, device=dev), dmeans3D=torch.empty(N, 3, device=dev), dopacity=torch.empty(N, device=dev),
              dshs=torch.empty(N, K, 3, device=dev), dscales=torch.empty(N, 3, device=dev), drotations=torch.empty(N, 4, device=dev))
    grads = L.TgsGrads(**{k: v.data_ptr() for k, v in gr.items()})
    ptrs = (C.c_void_p * 2)(state[0]["sg"].data_ptr(), state[1]["sg"].data_ptr())
    rows = (C.c_int32 * 4)(*[int(v) for b in bands for v in b])
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for it in range(3):
        L.check(lib.tgs_backward_preprocess_gather(C.byref(s0["st"]), C.byref(s0["gs"]), C.byref(s0["saved"]),
                                                   C.c_void_p(s0["radii"].data_ptr()), ptrs, rows, 2, C.byref(grads), stream), "gather")
    torch.cuda.synchronize(dev)
    # reference: summed buffers through the plain entry point
    rows0, rows1 = state[0]["sg"][: 10 * N].view(N, 10), state[1]["sg"][: 10 * N].view(N, 10)
    total = rows0 + rows1.to(dev)
    s0["st"].contrib_flags = 0                       # the summed buffer is plain [N,10]
    gr2 = {k: torch.empty_like(v) for k, v in gr.items()}
    grads2 = L.TgsGrads(**{k: v.data_ptr() for k, v in gr2.items()})
    L.check(lib.tgs_backward_preprocess(C.byref(s0["st"]), C.byref(s0["gs"]), C.byref(s0["saved"]), C.c_void_p(s0["radii"].data_ptr()),
                                        C.c_void_p(total.data_ptr()), C.byref(grads2), stream), "plain")
    torch.cuda.synchronize(dev)
    same = all(torch.equal(gr[k], gr2[k]) for k in gr)
    remote_rows = int((rows1.abs().sum(1) > 0).sum())
    print(f"contrib_flags={int(FLAGS)}: peer gather over NVLink == summed buffers: {same}; rank 1 holds {remote_rows} non-zero rows "
          f"({remote_rows * 40 / 1e6:.1f} MB of 40-byte rows)")
