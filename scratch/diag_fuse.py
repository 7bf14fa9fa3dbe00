import sys, os, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import touchgs_b200 as T
from oracle import fusion_oracle as F
Z = np.load(os.path.join(R, "tests", "golden", "fusion_reference.npz"))
sc = Z["case0_scalars"]
ins = {k: Z[f"case0_in_{k}"] for k in ("touch", "vision", "touch_sigma")}
dev = torch.device("cuda:0")
t = {k: torch.from_numpy(v.copy()).to(dev) for k, v in ins.items()}
print("roundtrip equal:", all(np.array_equal(t[k].cpu().numpy(), ins[k]) for k in ins), t["touch"].dtype, t["touch"].data_ptr() % 16)
got = T.touch_inputs.fuse_touch_vision(t["touch"], t["vision"], t["touch_sigma"], float(sc[0]), float(sc[1]), float(sc[2]), bool(sc[3]), 1.0)
torch.cuda.synchronize()
for k, g in zip(("vision_aligned", "ds_gs", "fused", "fused_sigma"), got[:4]):
    a = g.cpu().numpy(); ref = Z[f"case0_out_{k}"]
    bad = np.argwhere(a != ref)
    print(k, "mismatch", len(bad), "of", a.size, "flat%4 hist", np.bincount((bad[:, 0] * a.shape[1] + bad[:, 1]) % 4, minlength=4) if len(bad) else None)
    for b in bad[:6]:
        b = tuple(b); print("   ", b, "got", int(a[b]), "ref", int(ref[b]), "in t/v/s", int(ins["touch"][b]), int(ins["vision"][b]), int(ins["touch_sigma"][b]))
