"""CPU oracle (numpy, float64) of the per-pixel part of the reference's touch/vision depth fusion --
SURVEY.md §8(f) row N2, the data-format row feeding the hot path's touch target and weight.

STATUS: TEST INFRASTRUCTURE ONLY -- **PARITY PINNED**: unlike the rasterizer, this code IS in the
reference tree, so every function below restates reference lines and is checked bit-exactly (uint16
outputs) against golden vectors produced by running the reference's own functions
(``tests/golden/make_fusion_golden.py`` imports ``/root/reference/utils/fuse_touch_vision.py``).

Out of scope here (stays on the CPU, once per image): the two L-BFGS-B fits of (scale, offset)
(reference ``utils/fuse_touch_vision.py:41-74,285,301``); their results enter as scalars.
"""
from __future__ import annotations

import numpy as np

__all__ = ["decode_mm", "align_apply", "vision_sigma", "fuse_with_uncertainty", "encode_mm",
           "training_decode", "pipeline"]


def decode_mm(img_u16: np.ndarray) -> np.ndarray:
    """uint16 millimetres -> float64 metres (reference utils/fuse_touch_vision.py:270-276: ``img / 1000``)."""
    return img_u16 / 1000


def align_apply(vision, touch, scale, offset, offset2, is_real_world=True):
    """reference utils/fuse_touch_vision.py:288-306 (align_vision_depth) with the fitted scalars given.
    Returns (ds_gs_visual_depth, aligned vision depth)."""
    v = (scale * vision) + offset                                      # :288
    ds_gs = np.copy(v)                                                  # :291
    diff = v - touch                                                    # :294
    diff[diff > 3] = 0                                                  # :295
    touch_to_align = touch * (diff > 0) if is_real_world else touch     # :297
    mask = touch_to_align > 0                                           # :298
    v[mask] = v[mask] + offset2                                         # :304
    v = np.clip(v, a_min=0, a_max=None)                                 # :306
    return ds_gs, v


def vision_sigma(vision_aligned):
    """reference utils/fuse_touch_vision.py:310-313 calling
    utils/create_uncertainty_from_depth.py:9-58 with edge_weight=0, distance_uncertainty_weight=0.05,
    proximity_weight=0, depth_difference_weight=0: every term except ``dense_depth * 0.05`` (:21) is
    multiplied by an exact 0 (as long as the sparse grounded map has both zero and non-zero pixels, which
    the 1 % sparsification at :353 guarantees), then clip [0,10] (:312) and +5 (:313)."""
    return np.clip((vision_aligned ** 1) * 0.05, a_min=0, a_max=10) + 5


def fuse_with_uncertainty(touch, vision, touch_sigma, vision_sigma_):
    """reference utils/fuse_touch_vision.py:76-202 (fuse_depth_maps_with_uncertainty), float64."""
    with np.errstate(divide="ignore", invalid="ignore"):
        mask = touch_sigma > 0                                          # :109
        rv = 1 / vision_sigma_                                          # :116
        rt = 1 / touch_sigma                                            # :117
        rt[np.isinf(rt)] = 0                                            # :120
        rv[np.isinf(rv)] = 0                                            # :121
        sigma = 1 / (rt + rv)                                           # :124
        sigma[np.isinf(sigma)] = 0                                      # :126
        mu_t = (touch * mask) / touch_sigma                             # :136,140
        mu_t[np.isnan(mu_t)] = 0                                        # :141
        mu_v = vision / vision_sigma_                                   # :143
        mu_v[np.isnan(mu_v)] = 0                                        # :144
        fused = sigma * (mu_t + mu_v)                                   # :146
    return fused, sigma


def encode_mm(img):
    """reference utils/fuse_touch_vision.py:373-376 (save): ``(img * 1000).astype(np.uint16)``."""
    return (img * 1000).astype(np.uint16)


def training_decode(depth_u16, sigma_u16, scene_scale=1.0):
    """What the trainer feeds the operator: target = mm * 1e-3 * scene scale (reference
    legacy/dataparser_tactile.py:65-66,229-235: depth_unit_scale_factor = 1e-3 times the pose scale
    factor), 0 = invalid; weight = 1 / sigma (0 where sigma == 0).  fp32 tensors."""
    target = (depth_u16.astype(np.float64) * (1e-3 * scene_scale)).astype(np.float32)
    sig = sigma_u16.astype(np.float64) / 1000
    with np.errstate(divide="ignore"):
        w = np.where(sig > 0, 1.0 / sig, 0.0)
    return target, w.astype(np.float32)


def pipeline(touch_u16, vision_u16, touch_sigma_u16, scale, offset, offset2, is_real_world=True, scene_scale=1.0):
    """Everything per-pixel of reference fuse_vision_and_touch (:317-370) after the two fits."""
    touch, vision, tsig = decode_mm(touch_u16), decode_mm(vision_u16), decode_mm(touch_sigma_u16)
    ds_gs, v = align_apply(vision, touch, scale, offset, offset2, is_real_world)
    vs = vision_sigma(v)
    fused, sigma = fuse_with_uncertainty(touch, v, tsig, vs)
    fused = np.clip(fused, a_min=0, a_max=None)                         # :360
    sigma = np.clip(sigma, a_min=0, a_max=10)                           # :361
    out = dict(vision_aligned=encode_mm(v), ds_gs=encode_mm(ds_gs), fused=encode_mm(fused),
               fused_sigma=encode_mm(sigma))
    out["target"], out["weight"] = training_decode(out["fused"], out["fused_sigma"], scene_scale)
    return out
