"""GPU version of the per-pixel part of the reference's touch / vision depth fusion (SURVEY.md §8(f)
row N2): uint16-mm images in, the fused uint16-mm depth + uncertainty the reference would write to
disk out, plus the fp32 ``touch_depth`` / ``touch_weight`` tensors the rasterizer consumes.

Mirrors reference ``utils/fuse_touch_vision.py:317-370`` for one image, except the two L-BFGS-B fits
(``:285,:301``) whose results (scale, offset, offset2) the caller passes in.  Results are bit-identical to
the reference's PNG bytes (float64 arithmetic in the reference's operation order).
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple

import torch

from . import _lib as L


class FusedTouch(NamedTuple):
    vision_aligned_mm: torch.Tensor   # uint16 [H,W]  (reference output_dir/*.png)
    ds_gs_mm: torch.Tensor            # uint16 [H,W]  (reference output_dir_baseline/*.png)
    fused_mm: torch.Tensor            # uint16 [H,W]  (reference fused_output_dir/*.png)
    fused_sigma_mm: torch.Tensor      # uint16 [H,W]  (reference fused_output_dir_uncertainty/*.png)
    target: torch.Tensor              # fp32  [H,W]  -> GaussianRasterizer(touch_depth=...)
    weight: torch.Tensor              # fp32  [H,W]  -> GaussianRasterizer(touch_weight=...)


def _u16(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.device.type != "cuda":
        raise RuntimeError("touch_inputs.fuse_touch_vision is CUDA-only (no CPU fallback)")
    if t.dtype not in (torch.uint16, torch.int16):
        raise ValueError(f"{name} must be uint16 millimetres, got {t.dtype}")
    return t.contiguous()


def fuse_touch_vision(touch_mm, vision_mm, touch_sigma_mm, scale: float, offset: float, offset2: float,
                      is_real_world: bool = True, scene_scale: float = 1.0) -> FusedTouch:
    lib = L.load()
    touch_mm, vision_mm, touch_sigma_mm = (_u16(t, n) for t, n in
                                           ((touch_mm, "touch_mm"), (vision_mm, "vision_mm"), (touch_sigma_mm, "touch_sigma_mm")))
    if not (touch_mm.shape == vision_mm.shape == touch_sigma_mm.shape):
        raise ValueError("Depth maps must have the same shape.")     # reference utils/fuse_touch_vision.py:78
    dev, shape, n = touch_mm.device, touch_mm.shape, touch_mm.numel()
    with torch.cuda.device(dev):
        outs = [torch.empty(shape, dtype=torch.uint16, device=dev) for _ in range(4)]
        target = torch.empty(shape, dtype=torch.float32, device=dev)
        weight = torch.empty(shape, dtype=torch.float32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        L.check(lib.tgs_fuse_touch_vision(p(touch_mm), p(vision_mm), p(touch_sigma_mm), n, float(scale), float(offset),
                                          float(offset2), int(bool(is_real_world)), float(scene_scale), *(p(o) for o in outs),
                                          p(target), p(weight), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                "tgs_fuse_touch_vision")
    return FusedTouch(*outs, target, weight)
