"""How many Gaussians of the c3 step receive a non-zero screen-space gradient?  (python profiles/zero_rows.py [config])
The zero rows are what k_preprocess_bwd's ZSKIP shortcut and the contributor flags of the peer gather skip."""
import sys; sys.path.insert(0, '.')
import json, torch, bench, touchgs_b200 as T
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = dict(T.synth.CONFIGS[name])
dev = torch.device("cuda:0")
scene, params, batches, bg = bench.make_workload(cfg, cfg.get("N", 0), 2, dev, 0, None)
st = bench.Stepper(cfg, params, bg, dev, None, None)
out = {}
for ci, b in enumerate(batches):
    st.device_step(b)
    torch.cuda.synchronize()
    N = params["means3D"].shape[0]
    nz = (params["opacities"].grad.view(N, -1) != 0).any(1) | (params["shs"].grad.view(N, -1) != 0).any(1) | \
         (params["means3D"].grad.view(N, -1) != 0).any(1)
    out["cam%d" % ci] = {"N": N, "nonzero_rows": int(nz.sum()), "fraction": float(nz.float().mean())}
print(json.dumps(out))
