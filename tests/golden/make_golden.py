"""Generate tests/golden/*.npz from the CPU oracle.

The reference holds no golden vectors / known-answer tests for the rasterizer (it is not vendored:
reference .gitmodules:7-9), so these fixtures are produced by OUR oracle and pin it
(and, through the GPU parity tests, the CUDA path) against drift.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import touchgs_b200 as T  # noqa: E402
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_inputs():
    """A reduced c1 (BASELINE config #1 family): 400 Gaussians, 64x48, SH degree 1, L1 touch loss."""
    sc = T.synth.make_scene(400, 1, 0.02, 0.2, seed=11)
    cam = T.synth.look_at_camera(64, 48, (0.6, 0.4, -2.9))
    g = torch.Generator().manual_seed(99)
    grgb = torch.rand(3, 48, 64, generator=g) / (3 * 48 * 64)
    base = O.rasterize(sc.means3D, sc.opacities, _settings(sc, cam), shs=sc.shs, scales=sc.scales,
                       rotations=sc.rotations)
    target, weight = T.synth.make_touch_maps(base.depth[0] + 0.05, seed=11, n_patches=3, patch_radius=6)
    return sc, cam, grgb, target, weight


def _settings(sc, cam):
    return O.OracleSettings(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy,
                            torch.tensor([0.2, 0.1, 0.3]), 1.0, cam.viewmatrix, cam.projmatrix,
                            sc.sh_degree, cam.campos)


def golden_case():
    sc, cam, grgb, target, weight = golden_inputs()
    ins = {k: v.clone().requires_grad_(True) for k, v in
           dict(means3D=sc.means3D, scales=sc.scales, rotations=sc.rotations, opacities=sc.opacities, shs=sc.shs).items()}
    out = O.rasterize(ins["means3D"], ins["opacities"], _settings(sc, cam), shs=ins["shs"], scales=ins["scales"],
                      rotations=ins["rotations"], touch_depth=target, touch_weight=weight, depth_loss="l1",
                      depth_loss_mult=0.2, depth_normalize=True)
    ((out.color * grgb).sum() + out.touch_loss).backward()
    return out, {k: v.grad.detach() for k, v in ins.items()}


if __name__ == "__main__":
    out, grads = golden_case()
    sc, cam, grgb, target, weight = golden_inputs()
    np.savez_compressed(
        os.path.join(HERE, "oracle_c1_small.npz"),
        radii=out.radii.numpy(), keys=out.bins.keys.numpy(), vals=out.bins.vals.numpy(),
        ranges=out.bins.ranges.numpy(), n_contrib=out.img.n_contrib.numpy(),
        color=out.color.detach().numpy(), depth=out.depth.detach().numpy(), alpha=out.alpha.detach().numpy(),
        residual=out.residual.numpy(), target=target.numpy(), weight=weight.numpy(), grgb=grgb.numpy(),
        **{"grad_" + k: v.numpy() for k, v in grads.items()})
    print("wrote", os.path.join(HERE, "oracle_c1_small.npz"), "I =", out.bins.keys.numel())
