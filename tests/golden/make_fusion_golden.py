"""Generate tests/golden/fusion_reference.npz by RUNNING THE REFERENCE's own code
(/root/reference/utils/fuse_touch_vision.py) on seeded synthetic uint16 depth images.

Only runs in the build container (the reference tree is not on the GPU box); the fixture it writes
is committed.  matplotlib is absent here and is only used by the reference's viz=True branches, so it
is stubbed for the import.
"""
import os
import sys
import tempfile
import types

import cv2
import numpy as np

REF = "/root/reference/utils"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import fuse_touch_vision as f       # noqa: E402
    return f


def synth_images(h, w, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    true = 1.2 + 0.4 * np.sin(xx / 17.0) * np.cos(yy / 11.0) + 0.002 * xx          # metres
    grounded = np.where(rng.random((h, w)) < 0.9, true + rng.normal(0, 0.003, (h, w)), 0.0)   # holes
    vision = 0.7 * true + 0.35 + rng.normal(0, 0.02, (h, w))                        # scale/offset ambiguity
    touch = np.zeros((h, w)); tsig = np.zeros((h, w))
    for _ in range(4):                                                               # a few touch patches
        cy, cx, r = rng.integers(8, h - 8), rng.integers(8, w - 8), rng.integers(4, 9)
        disc = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        touch[disc] = true[disc] + rng.normal(0, 0.001, disc.sum())
        tsig[disc] = np.abs(rng.normal(0.004, 0.002, disc.sum())) + 0.001
    touch[0, :5] = true[0, :5]; tsig[0, :5] = 0.0          # touch depth with ZERO sigma (1/0 -> inf -> 0 branch)
    tsig[1, :5] = 0.003                                      # sigma without depth
    vision[2, :5] = 0.0                                      # holes in the vision map
    enc = lambda a: np.clip(a * 1000, 0, 65535).astype(np.uint16)
    return enc(grounded), enc(touch), enc(vision), enc(tsig)


def run_reference(f, grounded_u16, touch_u16, vision_u16, tsig_u16, is_real_world, seed):
    """Mirrors fuse_vision_and_touch (reference :317-370) for one image, calling the reference functions."""
    grounded, touch, vision, tsig = (a / 1000 for a in (grounded_u16, touch_u16, vision_u16, tsig_u16))   # :270-276
    np.random.seed(seed)                                               # create_sparse_depth_map uses np.random
    grounded = f.create_sparse_depth_map(grounded, keep_percentage=0.01)                     # :353
    # the two fits, recomputed exactly as align_vision_depth does, to record the scalars the kernel takes
    scale, offset = f.compute_scale_and_offset_best(grounded, vision, None, (0, None), (None, None))
    v1 = scale * vision + offset
    diff = v1 - touch; diff[diff > 3] = 0
    t2a = touch * (diff > 0) if is_real_world else touch
    _, offset2 = f.compute_scale_and_offset_best(t2a, v1, None, (1, 1), (None, None))
    ds_gs, v, vsig = f.align_vision_depth(grounded, touch, np.copy(vision), is_real_world=is_real_world)   # :355
    fused, sigma = f.fuse_depth_maps_with_uncertainty(touch, v, tsig, vsig, viz=False)      # :359
    fused = np.clip(fused, a_min=0, a_max=None)                                              # :360
    sigma = np.clip(sigma, a_min=0, a_max=10)                                                # :361
    with tempfile.TemporaryDirectory() as d:                                                 # the reference's own save()
        out, fo = os.path.join(d, "va"), os.path.join(d, "fu")
        for p in (out, out + "_baseline", fo, fo + "_uncertainty"):
            os.makedirs(p)
        f.save(out, fo, "0000", v, ds_gs, fused, sigma)                                      # :372-388
        rd = lambda p: cv2.imread(p, cv2.IMREAD_ANYDEPTH)
        res = dict(vision_aligned=rd(f"{out}/0000.png"), ds_gs=rd(f"{out}_baseline/0000.png"),
                   fused=rd(f"{fo}/0000.png"), fused_sigma=rd(f"{fo}_uncertainty/0000.png"))
    return dict(scale=float(scale), offset=float(offset), offset2=float(offset2), vsig_f64=vsig), res


if __name__ == "__main__":
    f = import_reference()
    save = {}
    for i, (h, w, real) in enumerate([(64, 96, True), (48, 80, False)]):
        g, t, v, s = synth_images(h, w, seed=100 + i)
        sc, res = run_reference(f, g, t, v, s, real, seed=7 + i)
        for k, a in dict(grounded=g, touch=t, vision=v, touch_sigma=s).items():
            save[f"case{i}_in_{k}"] = a
        for k, a in res.items():
            save[f"case{i}_out_{k}"] = a
        save[f"case{i}_scalars"] = np.array([sc["scale"], sc["offset"], sc["offset2"], float(real)])
        save[f"case{i}_vision_sigma_f64"] = sc["vsig_f64"]
        print(f"case {i}: scale {sc['scale']:.6f} offset {sc['offset']:.6f} offset2 {sc['offset2']:.6f}  "
              f"fused range {res['fused'].min()}..{res['fused'].max()} mm")
    np.savez_compressed(os.path.join(HERE, "fusion_reference.npz"), **save)
    print("wrote", os.path.join(HERE, "fusion_reference.npz"))
