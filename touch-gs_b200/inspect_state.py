"""Decode the saved byte buffers of a forward call into named tensors (parity tests / debugging).

Uses the layout introspection entry points of the C ABI (``tgs_geom_layout`` ...), so the tests
check exactly what the kernels wrote: radii, tile rects, tiles_touched, scan offsets, sort keys and
values (before and after the sort), tile ranges, packed records, final_T, n_contrib.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib as L
from .rasterizer import (GaussianRasterizationSettings, TouchOptions, _RasterizeGaussians, TILE)


def _view(buf: torch.Tensor, off: int, count: int, dtype: torch.dtype) -> torch.Tensor:
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    return buf[off:off + nbytes].view(dtype)


def forward_state(means3D, opacities, rs: GaussianRasterizationSettings, shs=None, colors_precomp=None,
                  scales=None, rotations=None, cov3D_precomp=None, opt: TouchOptions = None) -> Dict[str, torch.Tensor]:
    """Run the operator's forward (no grad) and return outputs + decoded internal state."""
    lib = L.load()
    opt = opt or TouchOptions()

    class Ctx:  # minimal stand-in for the autograd ctx
        def save_for_backward(self, *t): self.saved = t
        def set_materialize_grads(self, v): pass
        def mark_non_differentiable(self, *t): pass

    ctx = Ctx()
    e = torch.empty(0, device=means3D.device)
    with torch.no_grad():
        color, radii, depth, alpha, resid, _ = _RasterizeGaussians.forward(
            ctx, means3D, None, e if shs is None else shs, e if colors_precomp is None else colors_precomp,
            opacities, e if scales is None else scales, e if rotations is None else rotations,
            e if cov3D_precomp is None else cov3D_precomp, rs, opt)
    geom, binning, image = ctx.saved[-3], ctx.saved[-2], ctx.saved[-1]
    N = int(means3D.shape[0])
    I = ctx.num_rendered
    cap = getattr(ctx, 'capacity', I) or I
    H, W = rs.image_height, rs.image_width
    Tx, Ty = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    gl, bl, il = L.TgsGeomLayout(), L.TgsBinningLayout(), L.TgsImageLayout()
    lib.tgs_geom_layout(N, C.byref(gl))
    lib.tgs_binning_layout(cap, C.byref(bl))
    lib.tgs_image_layout(W, H, C.byref(il))
    rec = _view(geom, gl.records, N * 12, torch.float32).view(N, 12)
    rect = _view(geom, gl.rect, N * 2, torch.int32).view(N, 2)
    out = dict(
        color=color, radii=radii, depth=depth, alpha=alpha, residual=resid, num_rendered=I,
        xy=rec[:, 0:2], gdepth=rec[:, 2], gid=rec[:, 3].contiguous().view(torch.int32),
        conic=rec[:, 4:7], opacity=rec[:, 7], rgb=rec[:, 8:11],
        cov3D=_view(geom, gl.cov3D, N * 6, torch.float32).view(N, 6),
        tiles_touched=_view(geom, gl.tiles_touched, N, torch.int32),
        clamped=_view(geom, gl.clamped, N, torch.uint8),
        rect_min=torch.stack([rect[:, 0] & 0xFFFF, rect[:, 1] & 0xFFFF], -1),
        rect_max=torch.stack([(rect[:, 0] >> 16) & 0xFFFF, (rect[:, 1] >> 16) & 0xFFFF], -1),
        order=_view(geom, gl.order, N, torch.int32),
        vals=_view(binning, bl.vals_sorted, I, torch.int32),
        ranges=_view(image, il.ranges, Tx * Ty * 2, torch.int32).view(Tx * Ty, 2),
        num_rendered_device=int(_view(image, il.count, 2, torch.int32)[0].item()) & 0xFFFFFFFF,
        final_T=_view(image, il.final_T, H * W, torch.float32).view(H, W),
        n_contrib=_view(image, il.n_contrib, H * W, torch.int32).view(H, W),
        depth_raw=_view(image, il.depth_raw, H * W, torch.float32).view(H, W),
    )
    # The binning never materialises per-instance tile ids (binning.cu: positions come from counting rectangles):
    # the tile of sorted position j is the tile whose [start, end) range holds j.
    rg = out["ranges"].to(torch.int64) & 0xFFFFFFFF
    lens = (rg[:, 1] - rg[:, 0]).clamp_min(0)
    tile_of_range = torch.repeat_interleave(torch.arange(Tx * Ty, device=rg.device), lens)
    tid = torch.full((I,), -1, dtype=torch.int64, device=rg.device)
    if tile_of_range.numel() == I and I > 0:
        # ranges tile the list without gaps or overlaps iff starts are the exclusive scan of the lengths
        starts = torch.cumsum(lens, 0) - lens
        ok = bool((rg[lens > 0, 0] == starts[lens > 0]).all())
        if ok:
            tid = tile_of_range
    out["tile_ids"] = tid
    # no per-instance copy of the records exists any more (the compositing kernels gather them by id): the view a test
    # may want is simply the per-Gaussian table indexed by the sorted ids
    out["records"] = rec[out["vals"].long()]
    # the spec'd 64-bit sort key of every sorted instance: tile << 32 | bits(depth of its Gaussian)
    dbits = out["gdepth"].contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    out["keys"] = (out["tile_ids"] << 32) | dbits[out["vals"].long()]
    return out
